// H = 128 fast path of the forward chain: activations never leave tensor memory.
//
// One persistent CTA per SM keeps TWO 128-point tiles in flight.  TMEM (512 columns) holds, per tile slot g:
//   D   fp32 accumulator, 128 columns            [g*256      , +128)
//   A   next layer's operand, 16-bit packed      [g*256 + 128, +64)  (hi, or the fp16 copy)
//                                                [g*256 + 192, +64)  (lo of the bf16 split)
// so the MMA reads A from TMEM (tcgen05.mma with a TMEM A operand) and shared memory is left to the
// weights: a 12-deep ring of 16 KiB stages (three whole layers of bf16 hi/lo) streamed once per tile PAIR.
//   warps 0-7 / 8-15  activation warps of slot 0 / 1 (2 column halves x 4 TMEM lane quarters each): tcgen05.ld D ->
//                     +FiLM shift -> MUFU sin -> bf16 hi/lo (or fp16) -> tcgen05.st A; last layer: output head in registers.
//   warps 16, 17      MMA issuers, one per K slab of the A operand: a layer of a slot goes out as two halves, the first
//                     under the slot's own epilogue, in the order (slot 0, half 0) (slot 0, half 1) (slot 1, half 0) ...
//                     so that one slot's epilogue also runs under the other slot's MMAs (see the issuer section).
//   warp 18           one lane streams weight stages with cp.async.bulk (1-D TMA).
// Each weight stage is consumed by slot 0 then slot 1 before it is released (M = 256 rows per byte fetched from L2).
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "layout.cuh"
#include "ptx.cuh"
#include "tc_common.cuh"

namespace cnf {

constexpr int kTc2H = 128;
constexpr int kTc2EpiWarps = 16;                       // 2 tile slots x 2 column halves x 4 lane quarters
constexpr int kTc2Threads = (kTc2EpiWarps + 3) * 32;   // + two MMA issuer warps + weight producer warp
// (A 20th warp plus setmaxnreg re-partitioning -- control warpgroup 96 -> 64 registers, activation warps 96 -> 104; the CTA
// pool only holds what was released -- removed the stash variant's spills (-2.5 %) but cost the power-capped decode 2 %.)
constexpr int kTc2BwdThreads = kTc2Threads;
constexpr int kTc2SlotCols = 256;
constexpr int kTc2MaxLayers = 64;  // hidden layers the f16f8 scale table holds (host-checked)

struct Tc2SmemTail {
  float shift_s[2][2][kTc2H];   // [slot][layer parity][column]
  float y_part[2][kTileM][4];   // head partial sums of the upper column half, per slot
  float y_stage[2][kTileM * 4]; // the tile's decoded values [row][cout], staged for the vectorised store
  float w_first_s[4 * kTc2H];  // [cin][H]
  float w_out_s[4 * kTc2H];
  uint64_t b_full[kTcMaxStages];
  uint64_t b_empty[kTcMaxStages];
  uint64_t a_half[2];  // K slab 0 of the slot's A operand written and its accumulator drained (forward kernel)
  uint64_t a_q3[2];    // ... and columns 64-95 (the first half of K slab 1): forward kernel only
  uint64_t a_full[2];
  uint64_t d_full[2];
  uint64_t turn[2];  // issue token passed between the two MMA issuer warps
  uint32_t tmem_base;
  float inv_scale[kTc2MaxLayers];  // f16f8: 1/S_l per hidden layer (the accumulator holds S_l * W h)
};

__host__ __device__ constexpr size_t tc2_smem_bytes(int num_stages) {
  return 1024 + (size_t)num_stages * kStageBytes + sizeof(Tc2SmemTail);
}

// f16f8: see f16f8_operands16 (tc_common.cuh) for the operand formats.
__device__ __forceinline__ void tc2_store_a16_f16f8(uint32_t tmem_a, int c0, const float (&h)[16]) {
  uint32_t hi[8], lo8[4], a8[4];
  f16f8_operands16(h, hi, lo8, a8);
  const int slab = c0 >> 6, k0 = c0 & 63;
  ptx::tmem_st_32x32b_x8(tmem_a + c0 / 2, hi);
  ptx::tmem_st_32x32b_x4(tmem_a + 64 + slab * 32 + k0 / 4, lo8);
  ptx::tmem_st_32x32b_x4(tmem_a + 64 + slab * 32 + 16 + k0 / 4, a8);
}

template <int PREC, bool PACKED_MATH = true>
__device__ __forceinline__ void tc2_store_a(uint32_t tmem_a, int c0, const float (&h)[32]) {
  if constexpr (PREC == CNF_PREC_F16F8) {
    float h0[16], h1[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { h0[j] = h[j]; h1[j] = h[16 + j]; }
    tc2_store_a16_f16f8(tmem_a, c0, h0);
    tc2_store_a16_f16f8(tmem_a, c0 + 16, h1);
    return;
  }
  constexpr bool kSplit = (PREC == CNF_PREC_BF16X3);
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const float x0 = h[2 * e], x1 = h[2 * e + 1];
    if (kSplit) {
      hi[e] = ptx::pack_bf16x2(x0, x1);
      if (PACKED_MATH) {
        const float2 r = ptx::bf16x2_residual(hi[e], x0, x1);
        lo[e] = ptx::pack_bf16x2(r.x, r.y);
      } else {
        lo[e] = ptx::pack_bf16x2(x0 - ptx::bf16lo_to_f32(hi[e]), x1 - ptx::bf16hi_to_f32(hi[e]));
      }
    } else {
      hi[e] = ptx::pack_f16x2(x0, x1);
    }
  }
  ptx::tmem_st_32x32b_x16(tmem_a + c0 / 2, hi);
  if (kSplit) ptx::tmem_st_32x32b_x16(tmem_a + 64 + c0 / 2, lo);
}

// 16 activations (columns c0..c0+15 of this thread's row) -> 8 packed words per part -> tcgen05.st.x8.
// PACKED_MATH: residuals by fma.f32x2 (two columns per instruction); the stash variant of the kernel keeps scalar
// subtractions (its register pressure turns the 64-bit register pairs into spills: measured 6% slower).
template <int PREC, bool PACKED_MATH = true>
__device__ __forceinline__ void tc2_store_a16(uint32_t tmem_a, int c0, const float (&h)[16]) {
  if constexpr (PREC == CNF_PREC_F16F8) {
    tc2_store_a16_f16f8(tmem_a, c0, h);
    return;
  }
  constexpr bool kSplit = (PREC == CNF_PREC_BF16X3);
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float x0 = h[2 * e], x1 = h[2 * e + 1];
    if (kSplit) {
      hi[e] = ptx::pack_bf16x2_pinned(x0, x1);
      if (PACKED_MATH) {
        const float2 r = ptx::bf16x2_residual(hi[e], x0, x1);
        lo[e] = ptx::pack_bf16x2_pinned(r.x, r.y);
      } else {
        lo[e] = ptx::pack_bf16x2_pinned(x0 - ptx::bf16lo_to_f32(hi[e]), x1 - ptx::bf16hi_to_f32(hi[e]));
      }
    } else {
      hi[e] = ptx::pack_f16x2_pinned(x0, x1);
    }
  }
  ptx::tmem_st_32x32b_x8(tmem_a + c0 / 2, hi);
  if (kSplit) ptx::tmem_st_32x32b_x8(tmem_a + 64 + c0 / 2, lo);
}

// Columns of a warpgroup (column half hf), in the order it produces them: groups 0,1 lie in K slab 0 of the next
// layer's A operand (columns 0-63), so the first half of the next layer's MMAs can start when every warp is half-way
// through its epilogue; group 2 of the two warpgroups together covers columns 64-95 and group 3 columns 96-127, so
// the MMAs of K slab 1 go out in two quarters and only the last quarter trails the epilogue.
__device__ __forceinline__ constexpr int tc2_group_col(int hf, int c) {
  return c < 2 ? 32 * hf + 16 * c : 64 + 32 * (c - 2) + 16 * hf;
}
// Quarter issue of K slab 1 pays where the tensor pipe has slack: fp16 (8 MMAs per slot-layer) +8 %; f16f8 (16) +-0;
// bf16x3 (24, tensor-bound) -5 % -- measured on case1, 1,024 x 65,536.
__host__ __device__ constexpr bool tc2_quarter_issue(int prec) { return prec == CNF_PREC_FP16; }

// One hidden layer for this thread's row and its warpgroup's 64 columns.
// Software pipeline over four 16-column groups: the TMEM load of group c+2 and the sines (MUFU) of group c+1 are issued
// before the bf16 split / pack (ALU) of group c, so the XU and ALU pipes overlap inside the warp.
// Unless LAST, arrives on `a_half` once the whole accumulator row is in registers and K slab 0 of the A operand is
// written (the issuer may then overwrite D with the first 12 MMAs of the next layer), and on `a_full` at the end.
struct Tc2NoHook {
  __device__ __forceinline__ void operator()() const {}
};
// HF (the warpgroup's column half) is a template parameter: every column offset -- TMEM addresses, shift and head-weight
// loads, stash offsets -- is then an immediate; with a run-time hf the layer spent ~40 of its ~440 instructions per
// thread recomputing them (LOP3 / IADD3 / IMAD / LEA / R2UR).
template <int PREC, bool LAST, bool STASH, int HF, typename HalfHook>
__device__ __forceinline__ void tc2_hidden_layer_hf(uint32_t lane_base, uint32_t tmem_a,
                                                    const float* __restrict__ sbuf, const float* __restrict__ w_out_s,
                                                    int cout, float (&y)[4], __half* stash_l, uint64_t* a_half,
                                                    uint64_t* a_q3, uint64_t* a_full, float inv, HalfHook on_half) {
  constexpr int hf = HF;
  constexpr bool SCALED = (PREC == CNF_PREC_F16F8);
  uint32_t v[2][16];
  float hb[2][16];  // sines of group c in hb[c & 1] (the loop is fully unrolled: both are register arrays, no copies)
  ptx::tmem_ld_32x32b_x16(lane_base + tc2_group_col(hf, 0), v[0]);
  ptx::tmem_wait_ld();
  ptx::tmem_ld_32x32b_x16(lane_base + tc2_group_col(hf, 1), v[1]);
  tc_sines16<STASH, SCALED>(v[0], sbuf + tc2_group_col(hf, 0), hb[0],
                            STASH ? stash_l + (size_t)tc2_group_col(hf, 0) * kTileM : nullptr, inv);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int c0 = tc2_group_col(hf, c);
    float(&hcur)[16] = hb[c & 1];
    if (c + 1 < 4) {
      const int c1 = tc2_group_col(hf, c + 1);
      ptx::tmem_wait_ld();
      tc_sines16<STASH, SCALED>(v[(c + 1) & 1], sbuf + c1, hb[(c + 1) & 1],
                                STASH ? stash_l + (size_t)c1 * kTileM : nullptr, inv);
      if (c + 2 < 4) ptx::tmem_ld_32x32b_x16(lane_base + tc2_group_col(hf, c + 2), v[c & 1]);
    }
    if (!LAST) {
      tc2_store_a16<PREC, !STASH>(tmem_a, c0, hcur);
      if (c == 1) {
        ptx::tmem_wait_ld();  // group 3 (the last of D) is in registers
        ptx::tmem_wait_st();
        ptx::tc_fence_before();
        ptx::mbar_arrive(a_half);
        on_half();  // debug trace hook (empty in product builds)
      } else if (c == 2 && tc2_quarter_issue(PREC)) {
        ptx::tmem_wait_st();
        ptx::tc_fence_before();
        ptx::mbar_arrive(a_q3);
      } else if (c == 3) {
        ptx::tmem_wait_st();
        ptx::tc_fence_before();
        ptx::mbar_arrive(a_full);
      }
    } else {
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        if (o >= cout) continue;
        if (!STASH) {
          float2 acc = make_float2(y[o], 0.f);  // even / odd columns, two FMAs per instruction (fma.f32x2)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 w4 = *reinterpret_cast<const float4*>(w_out_s + o * kTc2H + c0 + q * 4);
            acc = __ffma2_rn(make_float2(w4.x, w4.y), make_float2(hcur[q * 4 + 0], hcur[q * 4 + 1]), acc);
            acc = __ffma2_rn(make_float2(w4.z, w4.w), make_float2(hcur[q * 4 + 2], hcur[q * 4 + 3]), acc);
          }
          y[o] = acc.x + acc.y;
        } else {  // scalar FMAs in the same even / odd order: bit-identical to the packed variant
          float acc_e = y[o], acc_o = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 w4 = *reinterpret_cast<const float4*>(w_out_s + o * kTc2H + c0 + q * 4);
            acc_e = fmaf(w4.x, hcur[q * 4 + 0], acc_e);
            acc_o = fmaf(w4.y, hcur[q * 4 + 1], acc_o);
            acc_e = fmaf(w4.z, hcur[q * 4 + 2], acc_e);
            acc_o = fmaf(w4.w, hcur[q * 4 + 3], acc_o);
          }
          y[o] = acc_e + acc_o;
        }
      }
    }
  }
}

template <int PREC, bool LAST, bool STASH, typename HalfHook = Tc2NoHook>
__device__ __forceinline__ void tc2_hidden_layer(uint32_t lane_base, uint32_t tmem_a, int hf,
                                                 const float* __restrict__ sbuf, const float* __restrict__ w_out_s,
                                                 int cout, float (&y)[4], __half* stash_l, uint64_t* a_half,
                                                 uint64_t* a_q3, uint64_t* a_full, float inv = 1.f,
                                                 HalfHook on_half = HalfHook()) {
  if (hf == 0)  // warp-uniform
    tc2_hidden_layer_hf<PREC, LAST, STASH, 0>(lane_base, tmem_a, sbuf, w_out_s, cout, y, stash_l, a_half, a_q3, a_full, inv,
                                              on_half);
  else
    tc2_hidden_layer_hf<PREC, LAST, STASH, 1>(lane_base, tmem_a, sbuf, w_out_s, cout, y, stash_l, a_half, a_q3, a_full, inv,
                                              on_half);
}

template <int PREC, bool STASH, bool PACKED>
__global__ void __launch_bounds__(kTc2Threads, 1) tc2_forward_kernel(cnf_dims d, const uint8_t* __restrict__ packed,
                                                                     const float* __restrict__ coords,
                                                                     int64_t coord_frame_stride,
                                                                     const float* __restrict__ shift,
                                                                     OutTargets outs, __half* __restrict__ stash,
                                                                     LossArgs loss, int64_t T, int64_t P,
                                                                     int num_stages) {
  constexpr int pack_rows = PACKED ? 1 : 0;
  constexpr int H = kTc2H;
  constexpr bool kSplit = (PREC == CNF_PREC_BF16X3);
  constexpr bool kF8 = (PREC == CNF_PREC_F16F8);
  constexpr int kParts = (kSplit || kF8) ? 2 : 1;
  constexpr int kSPL = (H / kSlabK) * kParts;  // stages per layer: 4 (bf16 hi/lo, or fp16 + fp8 stage) or 2 (fp16)
  constexpr uint32_t kIdesc = ptx::make_idesc_f16(kSplit ? 1u : 0u, kTileM, H);
  // f16f8: the fp8 stage's first two K=32 MMAs multiply e5m2(a_lo) by e4m3(S w), the last two e5m2(a) by e4m3(S w_lo)
  [[maybe_unused]] constexpr uint32_t kIdescF8 = ptx::make_idesc_f8(ptx::kF8E5M2, ptx::kF8E4M3, kTileM, H);
  constexpr int kMmaWarp = kTc2EpiWarps;  // warps kMmaWarp, kMmaWarp+1: MMA issuers of slot 0, 1; then the producer

  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  Tc2SmemTail* tail = reinterpret_cast<Tc2SmemTail*>(ring + (size_t)num_stages * kStageBytes);

  const PackedLayout lay = make_layout(d);
  const int nl = d.nl, cin = d.cin, cout = d.cout;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t PB = (P + kTileM - 1) / kTileM;
  const int64_t tiles = tc_num_tiles(T, P, pack_rows);
  const int64_t pairs = (tiles + 1) / 2;
  const int64_t SH = (int64_t)(nl + 1) * H;

  if (threadIdx.x == 0) {
    for (int s = 0; s < num_stages; ++s) {
      ptx::mbar_init(&tail->b_full[s], 1);
      ptx::mbar_init(&tail->b_empty[s], 2);  // released by the commits of both slots' issuers
    }
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(&tail->a_half[g], 256);
      ptx::mbar_init(&tail->a_q3[g], 256);
      ptx::mbar_init(&tail->a_full[g], 256);
      ptx::mbar_init(&tail->d_full[g], 2);  // one commit per issuer warp (K slab 0, K slab 1)
      ptx::mbar_init(&tail->turn[g], 1);
    }
    ptx::fence_mbar_init();
  }
  {  // small fp32 operands of layer 0 and of the head, once per CTA
    const float* w_first = reinterpret_cast<const float*>(packed + lay.w_first);
    const float* w_out = reinterpret_cast<const float*>(packed + lay.w_out);
    // coordinate-major in shared memory ([cin][H]): layer 0 reads four columns of one coordinate with one LDS.128
    for (int i = threadIdx.x; i < H * cin; i += kTc2Threads) tail->w_first_s[(i % cin) * H + i / cin] = w_first[i];
    for (int i = threadIdx.x; i < cout * H; i += kTc2Threads) tail->w_out_s[i] = w_out[i];
    if (kF8) {
      const float* sc = reinterpret_cast<const float*>(packed + lay.tc_scale);
      for (int i = threadIdx.x; i < nl && i < kTc2MaxLayers; i += kTc2Threads) tail->inv_scale[i] = sc[i];
    }
  }
  if (warp == kMmaWarp) {
    ptx::tmem_alloc(&tail->tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp < kTc2EpiWarps) {
    // ===================== activation warpgroups =====================
    // warpgroup (g, hf): tile slot g, columns [32*hf, +32) of K slab 0 and [64 + 32*hf, +32) of K slab 1 (see
    // tc2_group_col); thread = TMEM lane / query point `row`.
    const int g = warp / 8, hf = (warp / 4) & 1, wq = warp % 4;
    const int row = wq * 32 + lane;
    const int stage_col = 32 * hf + (row & 31) + 64 * ((row >> 5) & 1);  // column whose shift this thread stages (wq < 2)
    const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16) + g * kTc2SlotCols;
    const uint32_t tmem_a = lane_base + 128;
    const float* b_out = reinterpret_cast<const float*>(packed + lay.b_out);
    const uint32_t bar_wg = 1 + g * 2 + hf;  // named barrier of this warpgroup (128 threads)
    const uint32_t bar_slot = 5 + g;         // named barrier of the slot's two warpgroups (256 threads)
    uint32_t d_phase = 0;
    float loss_acc = 0.f;  // fused loss: this thread's share of sum r^2 over all its tiles (head warps only)
    CNF_TRACE_DECL;
    const bool tracer = (lane == 0);
    [[maybe_unused]] const int trole = 4 + warp;
    for (int64_t pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
      const int64_t tile = 2 * pair + g;
      if (tile >= tiles) continue;
      const RowMap rm = tc_row_map(tile, row, T, P, PB, pack_rows);
      const int64_t t = rm.t, p = rm.p;
      const bool valid = rm.valid;
      const float* sh = shift + t * SH;  // this row's frame (one frame per tile unless the tiles are packed)
      float x[4] = {0.f, 0.f, 0.f, 0.f};
      if (valid) {
        const float* cp = coords + t * coord_frame_stride + p * cin;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < cin) x[j] = cp[j];
      }
      __half* st_row = STASH ? stash + (size_t)tile * SH * kTileM + row * 8 : nullptr;  // see tc_common.cuh
      if (tracer) CNF_TRACE_EVENT(trole, 100);  // tile start
      // FiLM shifts: staged per layer in shared memory (one frame per tile), or read per row from global (packed tiles)
      if (!PACKED) {
        ptx::bar_sync(bar_wg, 128);  // everyone is done with the previous tile's shift buffers
        if (wq < 2) tail->shift_s[g][0][stage_col] = __ldg(sh + stage_col);
        ptx::bar_sync(bar_wg, 128);
      }

      // ---- layer 0 on CUDA cores (K = cin), always range-reduced.  The body is instantiated per cin: with a run-time
      // cin the predicated-off FMAs and loads of the missing coordinates still took issue slots, and the scalar
      // shared-memory loads made it 19 instructions per element against 7 in a hidden layer -- a fifth of the kernel's
      // instructions.  Same FMA order as before (shift, then coordinate 0, 1, ..): bit-identical results.
      auto layer0 = [&](auto cin_tag) {
      constexpr int CIN = decltype(cin_tag)::value;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int c0 = 32 * hf + 64 * half;
        float h[32];
        [[maybe_unused]] float cs0[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 s4 = PACKED ? __ldg(reinterpret_cast<const float4*>(sh + c0 + j))
                                   : *reinterpret_cast<const float4*>(&tail->shift_s[g][0][c0 + j]);
          float z[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
          for (int i = 0; i < CIN; ++i) {
            const float4 w4 = *reinterpret_cast<const float4*>(&tail->w_first_s[i * kTc2H + c0 + j]);
            z[0] = fmaf(w4.x, x[i], z[0]);
            z[1] = fmaf(w4.y, x[i], z[1]);
            z[2] = fmaf(w4.z, x[i], z[2]);
            z[3] = fmaf(w4.w, x[i], z[3]);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float r = ptx::reduce_2pi<PREC == CNF_PREC_BF16X3>(z[k]);
            h[j + k] = ptx::sin_approx(r);
            if (STASH) cs0[j + k] = ptx::cos_approx(r);
          }
        }
        tc2_store_a<PREC, !STASH>(tmem_a, c0, h);
        if (STASH) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            float c16[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) c16[j] = cs0[q * 16 + j];
            tc_stash16(st_row + (size_t)(c0 + q * 16) * kTileM, c16);
          }
        }
        ptx::tmem_wait_st();
        ptx::tc_fence_before();
        if (half == 1 && tc2_quarter_issue(PREC)) ptx::mbar_arrive(&tail->a_q3[g]);
        ptx::mbar_arrive(half == 0 ? &tail->a_half[g] : &tail->a_full[g]);  // K slab `half` of the A operand is in TMEM
      }
      };
      switch (cin) {
        case 1: layer0(std::integral_constant<int, 1>{}); break;
        case 2: layer0(std::integral_constant<int, 2>{}); break;
        case 3: layer0(std::integral_constant<int, 3>{}); break;
        default: layer0(std::integral_constant<int, 4>{}); break;
      }
      if (tracer) CNF_TRACE_EVENT(trole, 101);  // layer 0 done, a_full arrived

      float y[4] = {0.f, 0.f, 0.f, 0.f};
      // every hidden layer: wait for the layer's accumulator, activate, (re)write the A operand or run the head
      auto layer_prologue = [&](int l) {
        if (!PACKED) {  // stage the layer's FiLM shifts (packed tiles read them per row from global memory instead)
          float* stage = tail->shift_s[g][l & 1];
          if (wq < 2) stage[stage_col] = __ldg(sh + (size_t)l * H + stage_col);
          ptx::bar_sync(bar_wg, 128);
        }
        if (tracer) CNF_TRACE_EVENT(trole, 200 + l);  // start waiting for d_full
        // one warp of the slot polls the mbarrier; the other seven sleep on the hardware barrier (no issue slots)
        if (hf == 0 && wq == 0) ptx::mbar_wait(&tail->d_full[g], d_phase);
        d_phase ^= 1u;
        ptx::bar_sync(bar_slot, 256);
        ptx::tc_fence_after();
        if (tracer) CNF_TRACE_EVENT(trole, 300 + l);  // d_full observed
      };
      [[maybe_unused]] int cur_layer = 0;
      auto half_hook = [&]() {
        if (tracer) CNF_TRACE_EVENT(trole, 350 + cur_layer);  // a_half arrived
      };
#pragma unroll 1
      for (int l = 1; l < nl; ++l) {
        layer_prologue(l);
        cur_layer = l;
        // two call sites so that each sees a pointer of known address space (ld.shared vs ld.global, not generic)
        const float inv = kF8 ? tail->inv_scale[l - 1] : 1.f;
        if (!PACKED)
          tc2_hidden_layer<PREC, false, STASH>(lane_base, tmem_a, hf, tail->shift_s[g][l & 1], tail->w_out_s, cout, y,
                                               STASH ? st_row + (size_t)l * H * kTileM : nullptr, &tail->a_half[g],
                                               &tail->a_q3[g], &tail->a_full[g], inv, half_hook);
        else
          tc2_hidden_layer<PREC, false, STASH>(lane_base, tmem_a, hf, sh + (size_t)l * H, tail->w_out_s, cout, y,
                                               STASH ? st_row + (size_t)l * H * kTileM : nullptr, &tail->a_half[g],
                                               &tail->a_q3[g], &tail->a_full[g], inv, half_hook);
        if (tracer) CNF_TRACE_EVENT(trole, 400 + l);  // epilogue of layer l done
      }
      {
        layer_prologue(nl);
        const float inv = kF8 ? tail->inv_scale[nl - 1] : 1.f;
        if (!PACKED)
          tc2_hidden_layer<PREC, true, STASH>(lane_base, tmem_a, hf, tail->shift_s[g][nl & 1], tail->w_out_s, cout, y,
                                              STASH ? st_row + (size_t)nl * H * kTileM : nullptr, nullptr, nullptr, nullptr, inv);
        else
          tc2_hidden_layer<PREC, true, STASH>(lane_base, tmem_a, hf, sh + (size_t)nl * H, tail->w_out_s, cout, y,
                                              STASH ? st_row + (size_t)nl * H * kTileM : nullptr, nullptr, nullptr, nullptr, inv);
        if (tracer) CNF_TRACE_EVENT(trole, 400 + nl);
      }
      // ---- head: combine the two column halves, 12-byte store per point
      ptx::tc_fence_before();
      if (hf == 1) *reinterpret_cast<float4*>(tail->y_part[g][row]) = make_float4(y[0], y[1], y[2], y[3]);
      ptx::bar_sync(bar_slot, 256);
      if (hf == 0) {
        const float4 yp = *reinterpret_cast<const float4*>(tail->y_part[g][row]);
        float ys[4] = {y[0] + yp.x, y[1] + yp.y, y[2] + yp.z, y[3] + yp.w};
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (o < cout) ys[o] += __ldg(b_out + o);
        if (loss.y_meas != nullptr) loss_acc += tc_loss_row(loss, t, p, P, cout, valid, ys);
        if (outs.n == 1) {  // local target: 4*cout bytes per row straight from registers (L2 merges the sectors)
          if (valid && outs.ptr[0] != nullptr) {  // (the fused-loss entry point may skip the decoded field)
            float* op = outs.ptr[0] + (t * P + p) * cout;
#pragma unroll
            for (int o = 0; o < 4; ++o)
              if (o < cout) op[o] = ys[o];
          }
        } else {
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < cout) tail->y_stage[g][row * cout + o] = ys[o];
        }
      }
      if (outs.n > 1) {
        // fused all-gather: peer stores cross NVLink, where partial 32-byte sectors are expensive -- stage the tile and
        // write its (contiguous) range to every rank's buffer as full 16-byte vectors
        ptx::bar_sync(bar_slot, 256);
        int64_t q0;
        int nvalid;
        tc_tile_range(tile, T, P, PB, pack_rows, q0, nvalid);
        tc_store_tile(outs, tail->y_stage[g], q0, nvalid, cout, hf * 128 + row, 256);
      }
      __syncwarp();  // the store loops have lane-dependent trip counts: reconverge before warp-aligned instructions
    }
    if (loss.y_meas != nullptr && hf == 0) {  // one slot per head warp: no atomics, deterministic
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, off);
      if (lane == 0) loss.partials[(blockIdx.x * 8 + g * 4 + wq) % kLossPartials] = loss_acc;
    }
    ptx::tc_fence_before();
  } else if (warp < kMmaWarp + 2) {
    // ===================== MMA issuers (one warp per K slab) =====================
    // A layer of a slot is issued as two halves, one per K slab of the A operand: half 0 as soon as the slot's
    // activation warps have drained the accumulator and written slab 0 (a_half), half 1 when they are done (a_full),
    // so half of the layer's MMAs overlap the slot's own epilogue.  The halves go out in the fixed order
    //   (slot 0, half 0) (slot 0, half 1) (slot 1, half 0) (slot 1, half 1) (slot 0, half 0 of the next layer) ...
    // which keeps the two slots half a period apart: one slot's epilogue runs under the other slot's MMAs.
    // Warp `half` issues every half-`half`, and an issue token (turn[]) ping-pongs between the two warps.  A warp that
    // has issued MMAs is held back until the tensor pipe has taken them (their descriptors live in its uniform
    // registers), so consecutive halves must come from DIFFERENT warps: the next warp has already waited for its
    // operands and only needs the token.  Each warp polls with all lanes (converged control flow keeps the tcgen05
    // operands in uniform registers; a divergent single-lane loop cannot sustain the issue rate); one elected lane
    // issues.  The tensor pipe executes in issue order; d_full[g] counts one commit per warp, so it completes only
    // when both halves of the layer have finished.
    const int half = warp - kMmaWarp;
    constexpr int kSPH = kSPL / 2;  // stages per half: hi, lo of one K slab (one stage for fp16)
    constexpr bool kQuarter = tc2_quarter_issue(PREC);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t ring_addr = ptx::smem_u32(ring);
    uint32_t a_phase[2] = {0u, 0u};
    uint32_t turn_phase = half == 0 ? 1u : 0u;  // half 0 issues first (a fresh barrier passes a parity-1 wait)
    int slot0 = half * kSPH;  // ring slot of this warp's first stage of the current layer
    uint32_t ph0 = 0;         // its mbarrier parity
    CNF_TRACE_DECL;
    for (int64_t pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
      for (int l = 1; l <= nl; ++l) {
        {  // this K slab's weights (shared by both slots): they landed long ago, the ring is three layers deep
          int slot = slot0;
          uint32_t ph = ph0;
#pragma unroll
          for (int s = 0; s < kSPH; ++s) {
            ptx::mbar_wait(&tail->b_full[slot], ph);
            if (++slot >= num_stages) { slot = 0; ph ^= 1u; }
          }
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const bool mine = (2 * pair + g < tiles);  // an odd tile count leaves slot 1 idle in the last pair
          const uint32_t tmem_d = tmem_u + g * kTc2SlotCols;
          const uint32_t tmem_a = tmem_d + 128;
          // The K steps of this warp's stages: all four of each stage (q < 0), or those of the 32-column quarter q of
          // the slab (16-bit stages: steps 2q, 2q+1; the fp8 stage: step q of e5m2(a_lo) and step 2+q of e5m2(a)).
          auto issue = [&](int q, bool release) {
            int slot = slot0;
#pragma unroll
            for (int s = 0; s < kSPH; ++s) {
              const uint64_t b = ptx::make_desc_k_sw128(ring_addr + slot * kStageBytes);
              const int part = s % kParts;
              const bool f8part = kF8 && part == 1;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if (q >= 0 && (f8part ? (kk & 1) : (kk >> 1)) != q) continue;
                const uint32_t a_hi = tmem_a + (half * 4 + kk) * 8;  // 16 K elements = 8 packed columns
                if (part == 0) {
                  ptx::umma_f16_ts(tmem_d, a_hi, b + 2 * kk, kIdesc, (uint32_t)((half | kk) != 0));
                  if (kSplit) ptx::umma_f16_ts(tmem_d, a_hi + 64, b + 2 * kk, kIdesc, 1u);
                } else if (kF8) {  // 32 8-bit K elements = 8 packed columns of this K slab's fp8 operand
                  ptx::umma_f8_ts(tmem_d, tmem_a + 64 + half * 32 + kk * 8, b + 2 * kk, kIdescF8, 1u);
                } else {
                  ptx::umma_f16_ts(tmem_d, a_hi, b + 2 * kk, kIdesc, 1u);
                }
              }
              if (release) ptx::umma_commit(&tail->b_empty[slot]);  // (tracks every earlier MMA of this thread)
              if (++slot >= num_stages) slot = 0;
            }
          };
          if (mine) {
            if (lane == 0) CNF_TRACE_EVENT(2 + half, 1000 + 500 * g + l);  // start waiting for the A operand
            ptx::mbar_wait(half == 0 ? &tail->a_half[g] : kQuarter ? &tail->a_q3[g] : &tail->a_full[g], a_phase[g]);
            if (lane == 0) CNF_TRACE_EVENT(2 + half, 2000 + 500 * g + l);  // operands ready
          }
          ptx::mbar_wait(&tail->turn[half], turn_phase);
          turn_phase ^= 1u;
          ptx::tc_fence_after();
          if (kQuarter && half == 1 && mine) {
            // K slab 1 in two quarters: columns 64-95 now, 96-127 when the epilogue is complete.  This warp is held
            // back until the tensor pipe has taken the first quarter, well within the epilogue's last column group.
            if (ptx::elect_one()) issue(0, false);
            __syncwarp();
            ptx::mbar_wait(&tail->a_full[g], a_phase[g]);
            ptx::tc_fence_after();
          }
          if (mine) a_phase[g] ^= 1u;
          if (ptx::elect_one()) {
            if (mine) {
              issue(kQuarter && half == 1 ? 1 : -1, true);
              ptx::umma_commit(&tail->d_full[g]);
            } else {
              int slot = slot0;
#pragma unroll
              for (int s = 0; s < kSPH; ++s) {
                ptx::mbar_arrive(&tail->b_empty[slot]);  // idle slot: still release its share of the stage
                if (++slot >= num_stages) slot = 0;
              }
            }
            ptx::mbar_arrive(&tail->turn[half ^ 1]);
          }
          __syncwarp();
          if (mine && lane == 0) CNF_TRACE_EVENT(2 + half, 3000 + 500 * g + l);  // this half issued + committed
        }
        slot0 += kSPL;
        if (slot0 >= num_stages) { slot0 -= num_stages; ph0 ^= 1u; }
      }
    }
  } else if (warp == kMmaWarp + 2) {
    // ===================== weight producer =====================
    if (lane == 0) {
      const uint8_t* wsrc = packed + (kSplit ? lay.tc_fwd_x3 : kF8 ? lay.tc_fwd_f8 : lay.tc_fwd_h);
      int slot = 0;
      uint32_t phase = 0;
      for (int64_t pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
        for (int l = 0; l < nl; ++l) {
          const uint8_t* src = wsrc + (size_t)l * kSPL * kStageBytes;
          for (int s = 0; s < kSPL; ++s) {
            ptx::mbar_wait(&tail->b_empty[slot], phase ^ 1u);
            ptx::mbar_arrive_expect_tx(&tail->b_full[slot], kStageBytes);
            ptx::bulk_g2s(ring + (size_t)slot * kStageBytes, src + (size_t)s * kStageBytes, kStageBytes,
                          &tail->b_full[slot]);
            if (++slot == num_stages) { slot = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}


// ------------------------------------------------------------------------------------------ backward, H = 128
// Same skeleton as tc2_forward_kernel with the transposed weight image: per tile slot the A operand in TMEM is
// delta_l (bf16 hi/lo), the accumulator is delta_l * (w0 W_l), and the activation warpgroups multiply by the stashed
// cos of the layer below, reduce the tile's 128 points per column with a 16-shuffle transpose-reduce per 16 columns
// and add the column sums into gshift with one red.global per (warp, column).
template <bool PACKED>
__global__ void __launch_bounds__(kTc2BwdThreads, 1) tc2_backward_kernel(cnf_dims d, const uint8_t* __restrict__ packed,
                                                                      const float* __restrict__ gout,
                                                                      const __half* __restrict__ stash,
                                                                      float* __restrict__ gshift, int64_t T, int64_t P,
                                                                      int num_stages) {
  constexpr int pack_rows = PACKED ? 1 : 0;
  constexpr int H = kTc2H;
  constexpr int PREC = CNF_PREC_BF16X3;
  constexpr int kSPL = (H / kSlabK) * 2;
  constexpr uint32_t kIdesc = ptx::make_idesc_f16(1u, kTileM, H);
  constexpr int kMmaWarp = kTc2EpiWarps;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  Tc2SmemTail* tail = reinterpret_cast<Tc2SmemTail*>(ring + (size_t)num_stages * kStageBytes);

  const PackedLayout lay = make_layout(d);
  const int nl = d.nl, cout = d.cout;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t PB = (P + kTileM - 1) / kTileM;
  const int64_t tiles = tc_num_tiles(T, P, pack_rows);
  const int64_t pairs = (tiles + 1) / 2;
  const int64_t SH = (int64_t)(nl + 1) * H;

  if (threadIdx.x == 0) {
    for (int s = 0; s < num_stages; ++s) {
      ptx::mbar_init(&tail->b_full[s], 1);
      ptx::mbar_init(&tail->b_empty[s], 2);
    }
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(&tail->a_half[g], 256);
      ptx::mbar_init(&tail->a_full[g], 256);
      ptx::mbar_init(&tail->d_full[g], 2);  // one commit per issuer warp (K slab 0, K slab 1)
      ptx::mbar_init(&tail->turn[g], 1);
    }
    ptx::fence_mbar_init();
  }
  {
    const float* w_out = reinterpret_cast<const float*>(packed + lay.w_out);
    for (int i = threadIdx.x; i < cout * H; i += kTc2BwdThreads) tail->w_out_s[i] = w_out[i];
  }
  if (warp == kMmaWarp) {
    ptx::tmem_alloc(&tail->tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp < kTc2EpiWarps) {
    const int g = warp / 8, hf = (warp / 4) & 1, wq = warp % 4;
    const int row = wq * 32 + lane;
    const int col0 = 32 * hf;  // this warpgroup's columns: [32hf, +32) of K slab 0 and [64 + 32hf, +32) of K slab 1
    const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16) + g * kTc2SlotCols;
    const uint32_t tmem_a = lane_base + 128;
    const uint32_t bar_slot = 5 + g;
    uint32_t d_phase = 0;
    CNF_TRACE_DECL;
    const bool tracer = (lane == 0);
    [[maybe_unused]] const int trole = 4 + warp;
    for (int64_t pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
      const int64_t tile = 2 * pair + g;
      if (tile >= tiles) continue;
      const RowMap rm = tc_row_map(tile, row, T, P, PB, pack_rows);
      const int64_t t = rm.t, p = rm.p;
      const bool valid = rm.valid;
      const __half* st_row = stash + (size_t)tile * SH * kTileM + row * 8;  // tile-major stash, see tc_common.cuh
      float gy[4] = {0.f, 0.f, 0.f, 0.f};
      if (valid) {
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (o < cout) gy[o] = gout[(t * P + p) * cout + o];
      }
      // ---- seed: delta at the last sine layer, delta_nl = (gy . W_out) .* cos_nl
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int c0 = col0 + 64 * (c >> 1) + 16 * (c & 1);
        float cs[16], dl[16];
        tc_load_cos16(st_row + ((size_t)nl * H + c0) * kTileM, cs);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float gsum = 0.f;
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < cout) gsum = fmaf(gy[o], tail->w_out_s[o * H + c0 + j], gsum);
          dl[j] = gsum * cs[j];
        }
        tc2_store_a16<PREC>(tmem_a, c0, dl);
        if (c & 1) {  // K slab c/2 of the first operand is written
          ptx::tmem_wait_st();
          ptx::tc_fence_before();
          ptx::mbar_arrive(c == 1 ? &tail->a_half[g] : &tail->a_full[g]);
        }
        if (PACKED) tc_colsum16_rows(dl, lane, t, gshift + (size_t)nl * H + c0, SH);
        else tc_colsum16_to_global(dl, lane, gshift + t * SH + (size_t)nl * H + c0);
      }

#pragma unroll 1
      for (int l = nl; l >= 1; --l) {
        // the stashed cos of the layer below does not depend on the MMA: fetch this thread's 64 columns (8 chunks of
        // 16 bytes, each warp access 512 contiguous bytes) BEFORE waiting for the accumulator, so the latency is hidden
        uint4 cpk[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
          tc_load_cos_chunk(st_row + ((size_t)(l - 1) * H + col0 + 64 * (q >> 2) + 8 * (q & 3)) * kTileM, cpk[q]);
        if (tracer) CNF_TRACE_EVENT(trole, 200 + l);  // cos prefetch issued, start waiting for d_full
        if (hf == 0 && wq == 0) ptx::mbar_wait(&tail->d_full[g], d_phase);
        d_phase ^= 1u;
        ptx::bar_sync(bar_slot, 256);
        ptx::tc_fence_after();
        if (tracer) CNF_TRACE_EVENT(trole, 300 + l);  // d_full observed
        // Two pairs of 16-column groups: pair 0 = this warpgroup's 32 columns of K slab 0 of the next operand, pair 1 = of
        // K slab 1.  Working on a pair at a time gives the scheduler independent work (loads, multiplies, packs, one
        // 32-column transpose-reduce).  After pair 0 is stored, pair 1's accumulator values are drained into registers
        // and the thread arrives on a_half: the first half of the next layer's MMAs (which overwrite D) runs under pair
        // 0's column sums and all of pair 1; a_full is signalled before pair 1's column sums, which the MMAs do not need.
        uint32_t v0[16], v1[16];
        ptx::tmem_ld_32x32b_x16(lane_base + col0, v0);
        ptx::tmem_ld_32x32b_x16(lane_base + col0 + 16, v1);
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          const int c0 = col0 + 64 * pr;
          float dl[32];
          ptx::tmem_wait_ld();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t(&v)[16] = h == 0 ? v0 : v1;
            float cs[16];
            tc_unpack_cos16(cpk[4 * pr + 2 * h], cpk[4 * pr + 2 * h + 1], cs);
#pragma unroll
            for (int j = 0; j < 16; j += 2) {  // mul.f32x2: two columns per instruction
              const float2 m = __fmul2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])),
                                          make_float2(cs[j], cs[j + 1]));
              dl[16 * h + j] = m.x;
              dl[16 * h + j + 1] = m.y;
            }
          }
          if (l > 1) {
            float d0[16], d1[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { d0[j] = dl[j]; d1[j] = dl[16 + j]; }
            tc2_store_a16<PREC>(tmem_a, c0, d0);
            tc2_store_a16<PREC>(tmem_a, c0 + 16, d1);
          }
          // Packed tiles (sensor-sized shapes, latency-bound, more registers per row) skip the early hand-over: they
          // drain pair 1 only after pair 0's column sums and signal a_half together with a_full.
          if (pr == 0 && !PACKED) {  // drain the rest of the accumulator row (pair 1) into the registers pair 0 has vacated
            ptx::tmem_ld_32x32b_x16(lane_base + col0 + 64, v0);
            ptx::tmem_ld_32x32b_x16(lane_base + col0 + 64 + 16, v1);
          }
          if (l > 1 && (!PACKED || pr == 1)) {
            if (pr == 0) ptx::tmem_wait_ld();
            ptx::tmem_wait_st();
            ptx::tc_fence_before();
            if (PACKED) ptx::mbar_arrive(&tail->a_half[g]);
            ptx::mbar_arrive(pr == 0 ? &tail->a_half[g] : &tail->a_full[g]);
          }
          if (PACKED) {  // rows of several frames per warp
            tc_colsum32_rows(dl, lane, t, gshift + (size_t)(l - 1) * H + c0, SH);
            if (pr == 0) {
              ptx::tmem_ld_32x32b_x16(lane_base + col0 + 64, v0);
              ptx::tmem_ld_32x32b_x16(lane_base + col0 + 64 + 16, v1);
            }
          } else {
            tc_colsum32_to_global(dl, lane, gshift + t * SH + (size_t)(l - 1) * H + c0);
          }
        }
        ptx::tc_fence_before();
        if (tracer) CNF_TRACE_EVENT(trole, 400 + l);  // epilogue of layer l done
      }
    }
    ptx::tc_fence_before();
  } else if (warp < kMmaWarp + 2) {
    // MMA issuers, one warp per K slab: same half-layer scheme as tc2_forward_kernel (see there)
    const int half = warp - kMmaWarp;
    constexpr int kSPH = kSPL / 2;  // stages per half: hi, lo of one K slab
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t ring_addr = ptx::smem_u32(ring);
    uint32_t a_phase[2] = {0u, 0u};
    uint32_t turn_phase = half == 0 ? 1u : 0u;
    int slot0 = half * kSPH;
    uint32_t ph0 = 0;
    CNF_TRACE_DECL;
    for (int64_t pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
      for (int l = nl; l >= 1; --l) {
        {
          int slot = slot0;
          uint32_t ph = ph0;
#pragma unroll
          for (int s = 0; s < kSPH; ++s) {
            ptx::mbar_wait(&tail->b_full[slot], ph);
            if (++slot >= num_stages) { slot = 0; ph ^= 1u; }
          }
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const bool mine = (2 * pair + g < tiles);
          const uint32_t tmem_d = tmem_u + g * kTc2SlotCols;
          const uint32_t tmem_a = tmem_d + 128;
          if (mine) {
            ptx::mbar_wait(half == 0 ? &tail->a_half[g] : &tail->a_full[g], a_phase[g]);
            a_phase[g] ^= 1u;
            if (lane == 0) CNF_TRACE_EVENT(2 + half, 2000 + 500 * g + l);  // operands ready
          }
          ptx::mbar_wait(&tail->turn[half], turn_phase);
          turn_phase ^= 1u;
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            int slot = slot0;
#pragma unroll
            for (int s = 0; s < kSPH; ++s) {
              if (mine) {
                const uint64_t b = ptx::make_desc_k_sw128(ring_addr + slot * kStageBytes);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  const uint32_t a_hi = tmem_a + (half * 4 + kk) * 8;
                  if (s == 0) {
                    ptx::umma_f16_ts(tmem_d, a_hi, b + 2 * kk, kIdesc, (uint32_t)((half | kk) != 0));
                    ptx::umma_f16_ts(tmem_d, a_hi + 64, b + 2 * kk, kIdesc, 1u);
                  } else {
                    ptx::umma_f16_ts(tmem_d, a_hi, b + 2 * kk, kIdesc, 1u);
                  }
                }
                ptx::umma_commit(&tail->b_empty[slot]);
              } else {
                ptx::mbar_arrive(&tail->b_empty[slot]);
              }
              if (++slot >= num_stages) slot = 0;
            }
            if (mine) ptx::umma_commit(&tail->d_full[g]);
            ptx::mbar_arrive(&tail->turn[half ^ 1]);
          }
          __syncwarp();
          if (mine && lane == 0) CNF_TRACE_EVENT(2 + half, 3000 + 500 * g + l);  // this half issued + committed
        }
        slot0 += kSPL;
        if (slot0 >= num_stages) { slot0 -= num_stages; ph0 ^= 1u; }
      }
    }
  } else if (warp == kMmaWarp + 2) {
    if (lane == 0) {
      const uint8_t* wsrc = packed + lay.tc_bwd_x3;
      int slot = 0;
      uint32_t phase = 0;
      for (int64_t pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
        for (int l = nl - 1; l >= 0; --l) {
          const uint8_t* src = wsrc + (size_t)l * kSPL * kStageBytes;
          for (int s = 0; s < kSPL; ++s) {
            ptx::mbar_wait(&tail->b_empty[slot], phase ^ 1u);
            ptx::mbar_arrive_expect_tx(&tail->b_full[slot], kStageBytes);
            ptx::bulk_g2s(ring + (size_t)slot * kStageBytes, src + (size_t)s * kStageBytes, kStageBytes,
                          &tail->b_full[slot]);
            if (++slot == num_stages) { slot = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace cnf
