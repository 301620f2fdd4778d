// Pieces shared by the tensor-core kernels (tc_kernels.cuh: generic SS-mode, tc2_kernels.cuh: H = 128 TMEM-resident).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "layout.cuh"
#include "ptx.cuh"

namespace cnf {

constexpr int kTcMaxStages = 12;

// Where the decoded field goes: this rank's (T, P, cout) block inside up to 8 output buffers.  n == 1 is the plain
// case; n > 1 is the fused all-gather: the other pointers are the peers' buffers mapped over NVLink (each already
// offset to this rank's frame range), written with ordinary stores from the epilogue, so the gather costs no extra pass.
constexpr int kMaxOutTargets = 8;
struct OutTargets {
  float* ptr[kMaxOutTargets];
  int n;
  int vec_ok;  // every target has the same 16-byte phase (checked on the host): the tile store may use 16-byte vectors
};

// Fused DPS measurement loss (reference: guided_diffusion/condition_methods.py:30-31 with the mask multiply of
// measurements.py:91-97), evaluated in the forward kernel's head when y_meas != nullptr:
//     y_phys = ya*y + yb                        (y_normalizer.denormalize folded into an affine map)
//     r      = y_meas - mask * y_phys           (the residual whose Frobenius norm is the DPS distance; the operator
//                                                returns mask*phy_fields, measurements.py:97 -- callers that want
//                                                mask*(y_meas - y_phys) pass an already masked measurement)
//     gy     = -ya * mask * r                   (= ||r|| * d||r||/dy: the backward seed, normalised later by 1/||r||)
// Each head warp adds its sum of r^2 into its own slot of `partials` (no atomics, deterministic); a finalize kernel
// reduces the slots to ||r|| and 1/||r||.
constexpr int kLossPartials = 4096;  // slots: (CTA, head warp) -> CTA*8 + warp slot; covers 512 CTAs
struct LossArgs {
  const float* y_meas;  // (T, P, cout), nullptr = loss disabled
  const float* mask;    // nullptr = all ones; else per kind below
  int mask_kind;        // 1: (P) per point; 2: (P, cout); 3: (T, P, cout)
  float ya[4], yb[4];
  float* gy;            // (T, P, cout) out
  float* partials;      // [kLossPartials] out, pre-zeroed
};
// Residual, seed and squared error of one decoded point (this thread's row): ys = network output (cout values).
__device__ __forceinline__ float tc_loss_row(const LossArgs& la, int64_t t, int64_t p, int64_t P, int cout, bool valid,
                                             const float (&ys)[4]) {
  float sq = 0.f;
  if (valid) {
    const int64_t e = (t * P + p) * cout;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      if (o >= cout) continue;
      float m = 1.f;
      if (la.mask != nullptr)
        m = la.mask_kind == 1 ? __ldg(la.mask + p) : la.mask_kind == 2 ? __ldg(la.mask + p * cout + o) : __ldg(la.mask + e + o);
      const float r = __ldg(la.y_meas + e + o) - m * fmaf(la.ya[o], ys[o], la.yb[o]);
      la.gy[e + o] = -la.ya[o] * m * r;
      sq = fmaf(r, r, sq);
    }
  }
  return sq;
}

#ifdef CNF_TRACE
// Debug build only: per-role event trace of CTA 0 (role r writes (code, clock64) pairs at trace[r*8192 + 2*n]).
static __device__ unsigned long long* g_trace = nullptr;
__device__ __forceinline__ void trace_event(unsigned long long* buf, int role, int& n, unsigned long long code) {
  if (buf != nullptr && n < 4000) {  // two fire-and-forget stores: the pointer was fetched once, at kernel start
    buf[role * 8192 + 2 * n] = code;
    buf[role * 8192 + 2 * n + 1] = clock64();
    ++n;
  }
}
#define CNF_TRACE_DECL \
  int trace_n = 0;     \
  unsigned long long* trace_p = (blockIdx.x == 0) ? g_trace : nullptr
#define CNF_TRACE_EVENT(role, code) trace_event(trace_p, role, trace_n, code)
#else
#define CNF_TRACE_DECL
#define CNF_TRACE_EVENT(role, code)
#endif

// Backward stash (tensor-core precisions): fp16 cos of every sine argument, tile-major so that a warp's accesses are
// contiguous: element (tile, layer l, column n, row r) lives at
//     stash[ ((tile*(nl+1) + l)*H + n) / 8 ][ r ][ n % 8 ]      (16-byte chunk of 8 columns per row, 128 rows per chunk)
// i.e. with st_row = stash + tile*(nl+1)*H*128 + r*8, the chunk of columns [c, c+8) of layer l is st_row + (l*H + c)*128.
// Thirty-two rows x 16 bytes = 512 contiguous bytes per warp access (4 full lines instead of 32 partial ones), and rows
// past P have their own (padded) slots, so no lane ever skips an access.
constexpr int kStashChunkStride = 8 * kTileM;  // halfs between consecutive 8-column chunks of one row

// 16 fp16 cosines (columns c..c+15 of this thread's row) -> two 16-byte chunks; dst = st_row + (l*H + c)*128
__device__ __forceinline__ void tc_stash16(__half* dst, const float (&c)[16]) {
  uint32_t w[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) w[e] = ptx::pack_f16x2(c[2 * e], c[2 * e + 1]);
  *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
  *reinterpret_cast<uint4*>(dst + kStashChunkStride) = make_uint4(w[4], w[5], w[6], w[7]);
}

// Hidden layers feed the pre-activation straight to sin.approx / cos.approx: MUFU's own range reduction (an fp32
// multiply by 1/2pi, then the fractional turn) costs |z| * 2^-23 of absolute error, the same order as the fp32 rounding
// of z itself in the reference, whatever |z| (tests: test_large_film_shifts_keep_parity); only layer 0 is reduced
// explicitly.  SCALED: the accumulator holds S * (W h) (f16f8 precision, weights pre-scaled by a power of two per
// layer) and `inv` = 1/S is folded into the FiLM add as one fma.
template <bool STASH, bool SCALED = false>
__device__ __forceinline__ void tc_sines16(const uint32_t (&v)[16], const float* __restrict__ sbuf, float (&h)[16],
                                            __half* stash_dst, float inv = 1.f) {
  float cs[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 s4 = *reinterpret_cast<const float4*>(sbuf + q * 4);
    const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
    // accumulator + FiLM shift, two columns per instruction (add.f32x2: half the issue slots of scalar FADDs); the
    // stash variant keeps scalar adds (its register pressure turns the 64-bit pairs into spills: measured 6% slower)
    float zs[4];
    if (!STASH) {
      const float2 a01 = make_float2(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]));
      const float2 a23 = make_float2(__uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
      const float2 z01 = SCALED ? __ffma2_rn(a01, make_float2(inv, inv), make_float2(sv[0], sv[1]))
                                : __fadd2_rn(a01, make_float2(sv[0], sv[1]));
      const float2 z23 = SCALED ? __ffma2_rn(a23, make_float2(inv, inv), make_float2(sv[2], sv[3]))
                                : __fadd2_rn(a23, make_float2(sv[2], sv[3]));
      zs[0] = z01.x; zs[1] = z01.y; zs[2] = z23.x; zs[3] = z23.y;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        zs[e] = SCALED ? fmaf(__uint_as_float(v[q * 4 + e]), inv, sv[e]) : __uint_as_float(v[q * 4 + e]) + sv[e];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float r = zs[e];
      h[q * 4 + e] = ptx::sin_approx_pinned(r);
      if (STASH) cs[q * 4 + e] = ptx::cos_approx(r);
    }
  }
  if (STASH) tc_stash16(stash_dst, cs);  // stash_dst = st_row + (l*H + c)*128, always a valid slot
}

// f16f8 operand of 16 activations (columns c0..c0+15 of this thread's row, c0 % 16 == 0), TMEM columns relative to the
// slot's A area (128 columns): fp16 copy in [0,64) as for the other precisions; the 8-bit operand of K slab s = c0/64 in
// [64 + 32s, +32): bytes [0,64) = e5m2(a - fp16(a)) and bytes [64,128) = e5m2(a) of the slab's 64 columns, so that one
// K = 128 fp8 row pairs with the weight stage [e4m3(c S w) | e4m3(c S (w - fp16 w))]:
//     a w ~= a16 w16 + e5m2(a_lo) e4m3(w) + e5m2(a) e4m3(w_lo)
// The e5m2 values are the HIGH BYTES of fp16 words (same sign / exponent layout, mantissa truncated to 2 bits): one
// PRMT per four values on the ALU pipe.  cvt.rn.satfinite.e4m3x2 / e5m2x2 would round to nearest but run at MUFU-class
// throughput and doubled the epilogue's XU load (measured: slot epilogue 1.85 k -> 2.5 k clk); the mean truncation loss
// is instead folded into the packed fp8 weights as the constant kF8TruncComp (pack.cuh), which leaves the same
// zero-mean error as round-to-nearest (scripts/emulate_precision.py).
__device__ __forceinline__ void f16f8_operands16(const float (&h)[16], uint32_t (&hi)[8], uint32_t (&lo8)[4],
                                                 uint32_t (&a8)[4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float x0 = h[4 * q], x1 = h[4 * q + 1], x2 = h[4 * q + 2], x3 = h[4 * q + 3];
    hi[2 * q] = ptx::pack_f16x2_pinned(x0, x1);
    hi[2 * q + 1] = ptx::pack_f16x2_pinned(x2, x3);
    const float2 r01 = ptx::f16x2_residual(hi[2 * q], x0, x1);
    const float2 r23 = ptx::f16x2_residual(hi[2 * q + 1], x2, x3);
    lo8[q] = __byte_perm(ptx::pack_f16x2_pinned(r01.x, r01.y), ptx::pack_f16x2_pinned(r23.x, r23.y), 0x7531);
    a8[q] = __byte_perm(hi[2 * q], hi[2 * q + 1], 0x7531);
  }
}
#ifdef CNF_TRACE
#define CNF_CHECK(cond, what, a, b)                                                                             \
  do {                                                                                                          \
    if (!(cond)) {                                                                                              \
      printf("CNF_CHECK %s failed: %lld %lld (block %d thread %d)\n", what, (long long)(a), (long long)(b),     \
             (int)blockIdx.x, (int)threadIdx.x);                                                                \
      __trap();                                                                                                 \
    }                                                                                                           \
  } while (0)
#else
#define CNF_CHECK(cond, what, a, b)
#endif

__device__ __forceinline__ void tc_colsum16_to_global(float (&v)[16], int lane, float* dst) {
  // 4 halving rounds: after them lane L holds column (L >> 1) & 15 summed over the 16 lanes with the same bit 0
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool upper = (lane & (off * 2)) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float recv = __shfl_xor_sync(0xffffffffu, send, off * 2);
      v[i] = (upper ? v[i + off] : v[i]) + recv;
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  if ((lane & 1) == 0) atomicAdd(dst + (lane >> 1), v[0]);
}

// 16 stashed cosines of this thread's row (columns c..c+15): src = st_row + (l*H + c)*128
__device__ __forceinline__ void tc_load_cos16(const __half* src, float (&c)[16]) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(src + q * kStashChunkStride));
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&ws[e]));
      c[q * 8 + 2 * e] = f.x;
      c[q * 8 + 2 * e + 1] = f.y;
    }
  }
}

// Which (frame, point) a tile row decodes.  Frame-aligned tiles (default): tile = (frame, 128-point block), rows past P
// are padding.  Packed tiles (small or ragged P, e.g. the DPS sensor shapes with 10 points per frame): the T*P
// (frame, point) pairs are numbered consecutively and cut into 128-row tiles, so a tile may span several frames and only
// the very last tile has padding.  Padding rows are clamped to the last valid pair (safe addresses, results discarded).
struct RowMap {
  int64_t t, p;
  bool valid;
};
__device__ __forceinline__ RowMap tc_row_map(int64_t tile, int row, int64_t T, int64_t P, int64_t PB, int packed) {
  RowMap m;
  if (packed) {
    const int64_t R = T * P, q = tile * kTileM + row;
    m.valid = q < R;
    const int64_t qq = m.valid ? q : R - 1;
    m.t = qq / P;
    m.p = qq - m.t * P;
  } else {
    m.t = tile / PB;
    m.p = (tile - m.t * PB) * kTileM + row;
    m.valid = m.p < P;
    if (!m.valid) m.p = P - 1;
  }
  return m;
}
__host__ __device__ inline int64_t tc_num_tiles(int64_t T, int64_t P, int packed) {
  return packed ? (T * P + kTileM - 1) / kTileM : T * ((P + kTileM - 1) / kTileM);
}

// Column sums of a warp's 32 rows when the rows may belong to different frames (packed tiles): `t_lane` is
// non-decreasing with the lane.  One frame: the plain transpose-reduce; a few frames: one masked reduce per frame;
// many frames (tiny P): per-row atomics.  `dst0` = gshift + layer/column offset; frame f's row is dst0 + f*SH.
// 32 columns at once: five halving rounds (16 + 8 + 4 + 2 + 1 shuffles, the 16 of the first round independent of
// each other); afterwards lane L holds column L summed over the warp's 32 rows -> one red.global per lane.
__device__ __forceinline__ void tc_colsum32_to_global(float (&v)[32], int lane, float* dst) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float recv = __shfl_xor_sync(0xffffffffu, send, off);
      v[i] = (upper ? v[i + off] : v[i]) + recv;
    }
  }
  atomicAdd(dst + lane, v[0]);
}
__device__ __forceinline__ void tc_colsum16_rows(float (&v)[16], int lane, int64_t t_lane, float* dst0, int64_t SH) {
  const int64_t t_lo = __shfl_sync(0xffffffffu, t_lane, 0), t_hi = __shfl_sync(0xffffffffu, t_lane, 31);
  if (t_lo == t_hi) {
    tc_colsum16_to_global(v, lane, dst0 + t_lo * SH);
  } else if (t_hi - t_lo < 4) {
    for (int64_t f = t_lo; f <= t_hi; ++f) {
      float w[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) w[j] = (t_lane == f) ? v[j] : 0.f;
      tc_colsum16_to_global(w, lane, dst0 + f * SH);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) atomicAdd(dst0 + t_lane * SH + j, v[j]);
  }
}
__device__ __forceinline__ void tc_colsum32_rows(float (&v)[32], int lane, int64_t t_lane, float* dst0, int64_t SH) {
  const int64_t t_lo = __shfl_sync(0xffffffffu, t_lane, 0), t_hi = __shfl_sync(0xffffffffu, t_lane, 31);
  if (t_lo == t_hi) {
    tc_colsum32_to_global(v, lane, dst0 + t_lo * SH);
  } else if (t_hi - t_lo < 4) {
    for (int64_t f = t_lo; f <= t_hi; ++f) {
      float w[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) w[j] = (t_lane == f) ? v[j] : 0.f;
      tc_colsum32_to_global(w, lane, dst0 + f * SH);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) atomicAdd(dst0 + t_lane * SH + j, v[j]);
  }
}

// First (frame, point) pair index and number of valid rows of a tile; a tile's valid rows are always a prefix and map
// to CONSECUTIVE pairs q0, q0+1, ... (frame-aligned tiles stay inside one frame), i.e. to one contiguous range of `out`.
__device__ __forceinline__ void tc_tile_range(int64_t tile, int64_t T, int64_t P, int64_t PB, int pack_rows, int64_t& q0,
                                              int& nvalid) {
  if (pack_rows) {
    q0 = tile * kTileM;
    const int64_t left = T * P - q0;
    nvalid = left < kTileM ? (int)left : kTileM;
  } else {
    const int64_t t = tile / PB, p0 = (tile - t * PB) * kTileM;
    q0 = t * P + p0;
    const int64_t left = P - p0;
    nvalid = left < kTileM ? (int)left : kTileM;
  }
}

// Cooperative store of a tile's decoded values, staged in shared memory as ys[row*cout + o], to every output target:
// the range is contiguous, so it goes out as full 16-byte vectors (full 32-byte sectors on the wire -- what matters for
// the fused all-gather, whose peer stores cross NVLink) with a scalar path for unaligned ranges.
__device__ __forceinline__ void tc_store_tile(const OutTargets& outs, const float* ys, int64_t q0, int nvalid, int cout,
                                              int tid, int nthreads) {
  const int64_t e0 = q0 * cout;
  const int n = nvalid * cout;
  // vec_ok: all targets share target 0's 16-byte phase (host-checked), so target 0 decides the alignment
  if (outs.vec_ok && (reinterpret_cast<uintptr_t>(outs.ptr[0] + e0) & 15) == 0) {
    const int n4 = n / 4;
    for (int i = tid; i < n4; i += nthreads) {
      const float4 v = *reinterpret_cast<const float4*>(ys + 4 * i);
      for (int k = 0; k < outs.n; ++k) *reinterpret_cast<float4*>(outs.ptr[k] + e0 + 4 * i) = v;
    }
    for (int i = 4 * n4 + tid; i < n; i += nthreads)
      for (int k = 0; k < outs.n; ++k) outs.ptr[k][e0 + i] = ys[i];
  } else {
    for (int i = tid; i < n; i += nthreads)
      for (int k = 0; k < outs.n; ++k) outs.ptr[k][e0 + i] = ys[i];
  }
}

// Raw (still packed) variant for prefetching a whole column range before the accumulator is ready.
__device__ __forceinline__ void tc_load_cos_chunk(const __half* src, uint4& w) {
  w = __ldg(reinterpret_cast<const uint4*>(src));
}
__device__ __forceinline__ void tc_unpack_cos16(const uint4& w0, const uint4& w1, float (&c)[16]) {
  const uint32_t ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&ws[e]));
    c[2 * e] = f.x;
    c[2 * e + 1] = f.y;
  }
}

}  // namespace cnf
