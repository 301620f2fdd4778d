// Weight packing kernels: flat fp32 parameters (state_dict order) -> PackedLayout.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "layout.cuh"

namespace cnf {

// fp32 sections: w0 folding, stacking and the transposed copy.
__global__ void pack_fp32_kernel(cnf_dims d, const float* __restrict__ params, float w0, uint8_t* __restrict__ packed) {
  const PackedLayout lay = make_layout(d);
  const ParamOffsets po = make_param_offsets(d);
  const size_t H = d.H, L = d.L, nl = d.nl;
  float* w_first = reinterpret_cast<float*>(packed + lay.w_first);
  float* w_out = reinterpret_cast<float*>(packed + lay.w_out);
  float* b_out = reinterpret_cast<float*>(packed + lay.b_out);
  float* b_shift = reinterpret_cast<float*>(packed + lay.b_shift);
  float* v_cat = reinterpret_cast<float*>(packed + lay.v_cat);
  float* w_hid = reinterpret_cast<float*>(packed + lay.w_hid);
  float* w_hid_t = reinterpret_cast<float*>(packed + lay.w_hid_t);

  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = tid; i < H * d.cin; i += stride) w_first[i] = w0 * params[po.w_first + i];
  float* w_first_t = reinterpret_cast<float*>(packed + lay.w_first_t);
  for (size_t i = tid; i < 4 * H; i += stride) {
    const size_t c = i / H, n = i % H;
    w_first_t[i] = c < (size_t)d.cin ? w0 * params[po.w_first + n * d.cin + c] : 0.f;
  }
  for (size_t i = tid; i < (size_t)d.cout * H; i += stride) w_out[i] = params[po.w_out + i];
  for (size_t i = tid; i < (size_t)d.cout; i += stride) b_out[i] = params[po.b_out + i];
  for (size_t i = tid; i < (nl + 1) * H; i += stride) {
    const size_t l = i / H, n = i % H;
    const size_t src = (l == 0) ? po.b_first + n : po.hid0 + (l - 1) * (H * H + H) + H * H + n;
    b_shift[i] = w0 * params[src];
  }
  for (size_t i = tid; i < (nl + 1) * H * L; i += stride) v_cat[i] = w0 * params[po.v0 + i];
  for (size_t i = tid; i < nl * H * H; i += stride) {
    const size_t l = i / (H * H), rem = i % (H * H), n = rem / H, k = rem % H;
    const float w = w0 * params[po.hid0 + l * (H * H + H) + rem];
    w_hid[i] = w;
    w_hid_t[l * H * H + k * H + n] = w;
  }
}

// Tensor-core stage images.  One thread per 16-bit output element.
//   mode 0: forward  bf16 hi/lo   B[n][k] = w0*W[n][k]
//   mode 1: forward  fp16         B[n][k] = w0*W[n][k]
//   mode 2: backward bf16 hi/lo   B[n][k] = w0*W[k][n]
__global__ void pack_tc_kernel(cnf_dims d, const float* __restrict__ params, float w0, uint8_t* __restrict__ packed,
                               int mode) {
  const PackedLayout lay = make_layout(d);
  const ParamOffsets po = make_param_offsets(d);
  const size_t H = d.H, nl = d.nl;
  const int parts = (mode == 1) ? 1 : 2;
  const size_t spl = stages_per_layer(d.H, parts);
  const size_t elems_per_stage = kStageRows * kSlabK;
  const size_t total = nl * spl * elems_per_stage;
  uint8_t* base = packed + (mode == 0 ? lay.tc_fwd_x3 : mode == 1 ? lay.tc_fwd_h : lay.tc_bwd_x3);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t l = i / (spl * elems_per_stage);
    const size_t s = (i / elems_per_stage) % spl;
    const uint32_t e = (uint32_t)(i % elems_per_stage);
    const uint32_t r = e / kSlabK, kk = e % kSlabK;
    // stage order = consumption order of the MMA warps: K slab, then hi/lo part, then 128-row block (innermost, so the
    // row blocks of one (slab, part) are adjacent in the ring and can feed a single N=256 MMA)
    const size_t nblocks = H / kStageRows;
    const size_t nb = s % nblocks;
    const size_t part = (s / nblocks) % parts;
    const size_t ks = s / (nblocks * parts);
    const size_t n = nb * kStageRows + r;
    const size_t k = ks * kSlabK + kk;
    const size_t src = (mode == 2) ? (k * H + n) : (n * H + k);
    const float w = w0 * params[po.hid0 + l * (H * H + H) + src];
    uint16_t bits;
    if (mode == 1) {
      bits = __half_as_ushort(__float2half_rn(w));
    } else {
      const __nv_bfloat16 hi = __float2bfloat16_rn(w);
      if (part == 0) {
        bits = __bfloat16_as_ushort(hi);
      } else {
        bits = __bfloat16_as_ushort(__float2bfloat16_rn(w - __bfloat162float(hi)));
      }
    }
    uint8_t* dst = base + (l * spl + s) * (size_t)kStageBytes + sw128_offset(r, kk);
    *reinterpret_cast<uint16_t*>(dst) = bits;
  }
}

// f16f8 precision, step 1: per hidden layer the power-of-two scale S_l that brings max|w0*W_l| just under 224 (so that
// S*w fits e4m3's +-448 with headroom and fp16's range trivially); block l handles layer l.
__global__ void __launch_bounds__(256) pack_scale_kernel(cnf_dims d, const float* __restrict__ params, float w0,
                                                         uint8_t* __restrict__ packed) {
  __shared__ float red[8];
  const PackedLayout lay = make_layout(d);
  const ParamOffsets po = make_param_offsets(d);
  const size_t H = d.H, l = blockIdx.x;
  const float* w = params + po.hid0 + l * (H * H + H);
  float m = 0.f;
  for (size_t i = threadIdx.x; i < H * H; i += blockDim.x) m = fmaxf(m, fabsf(w0 * w[i]));
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    float S = 1.f;
    if (m > 0.f && isfinite(m)) {
      int e;
      frexpf(224.f / m, &e);  // 224/m = f * 2^e, f in [0.5, 1)  ->  2^(e-1) <= 224/m
      e = e - 1;
      e = e > 30 ? 30 : (e < -30 ? -30 : e);
      S = ldexpf(1.f, e);
    }
    float* sc = reinterpret_cast<float*>(packed + lay.tc_scale);
    sc[l] = 1.f / S;
    sc[d.nl + l] = S;
  }
}

// The kernels build their e5m2 activation operands by TRUNCATING fp16 words to their high byte (2 mantissa bits); for a
// uniformly distributed mantissa the least-squares factor that re-centres the truncated value is 1.0873, folded here
// into the fp8 weight images that multiply those operands (scripts/emulate_precision.py: same error as rounding).
constexpr float kF8TruncComp = 1.0873f;

// f16f8 precision, step 2: the stage images.  One thread per weight element.
__global__ void pack_tc_f8_kernel(cnf_dims d, const float* __restrict__ params, float w0, uint8_t* __restrict__ packed) {
  const PackedLayout lay = make_layout(d);
  const ParamOffsets po = make_param_offsets(d);
  const size_t H = d.H, nl = d.nl;
  const size_t nblocks = H / kStageRows;
  const size_t spl = stages_per_layer(d.H, 2);
  const float* sc = reinterpret_cast<const float*>(packed + lay.tc_scale);
  uint8_t* base = packed + lay.tc_fwd_f8;
  const size_t total = nl * H * H;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t l = i / (H * H), rem = i % (H * H), n = rem / H, k = rem % H;
    const float w = w0 * params[po.hid0 + l * (H * H + H) + rem] * sc[nl + l];  // exact: S is a power of two
    const __half h = __float2half_rn(w);
    const float lo = w - __half2float(h);
    const size_t ks = k / kSlabK, nb = n / kStageRows;
    const uint32_t r = (uint32_t)(n % kStageRows), kk = (uint32_t)(k % kSlabK);
    // stage order (as for the bf16 hi/lo image): K slab -> part -> 128-row block
    uint8_t* st16 = base + (l * spl + (ks * 2 + 0) * nblocks + nb) * (size_t)kStageBytes;
    uint8_t* st8 = base + (l * spl + (ks * 2 + 1) * nblocks + nb) * (size_t)kStageBytes;
    *reinterpret_cast<uint16_t*>(st16 + sw128_offset(r, kk)) = __half_as_ushort(h);
    st8[sw128_byte_offset(r, kk)] = (uint8_t)__nv_cvt_float_to_fp8(kF8TruncComp * w, __NV_SATFINITE, __NV_E4M3);
    st8[sw128_byte_offset(r, 64 + kk)] = (uint8_t)__nv_cvt_float_to_fp8(kF8TruncComp * lo, __NV_SATFINITE, __NV_E4M3);
  }
}

}  // namespace cnf
