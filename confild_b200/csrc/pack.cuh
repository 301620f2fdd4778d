// Weight packing kernels: flat fp32 parameters (state_dict order) -> PackedLayout.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

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

}  // namespace cnf
