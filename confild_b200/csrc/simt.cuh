// CUDA-core fp32 kernels: the FiLM-shift GEMMs (always used) and the CNF_PREC_FP32 decode /
// backward chain (exact-order GPU reference, any hidden width).
#pragma once
#include <cuda_runtime.h>

#include "layout.cuh"
#include "tc_common.cuh"

namespace cnf {

// ------------------------------------------------------------------------------------------
// C[m][n] (+)= scale * sum_k A[m][k] * Bop[k][n] (+ bias[n]);  B_NK: B stored [N][K] (row-dot), else [K][N].
// 64x64 output tile per 256-thread block, 4x4 outputs per thread, K staged 16 at a time through double-buffered shared
// memory (global loads of step i+1 in flight under the FMAs of step i; operands read back as 16-byte vectors).
// SPLITK: blockIdx.z owns the K range [z*kc, (z+1)*kc) and adds its partial tile to a pre-zeroed C with atomics -- the
// DPS shapes (K4: M = 64..384 frames, N = L <= 384, K = (nl+1)H up to 6,144) have only a handful of output tiles.
template <bool B_NK, bool SPLITK>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                        const float* __restrict__ bias, float* __restrict__ C,
                                                        int64_t M, int N, int K, int kc,
                                                        const float* __restrict__ scale) {
  constexpr int LD = 68;  // row stride in floats: 16-byte aligned rows, conflict-free transposed stores
  __shared__ __align__(16) float As[2][16][LD];
  __shared__ __align__(16) float Bs[2][16][LD];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int64_t m0 = (int64_t)blockIdx.y * 64;
  const int n0 = blockIdx.x * 64;
  const int k_begin = SPLITK ? (int)blockIdx.z * kc : 0;
  const int k_end = SPLITK ? min(K, k_begin + kc) : K;
  float acc[4][4] = {};
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      {
        const int m = e / 16, k = e % 16;
        ra[i] = (m0 + m < M && k0 + k < k_end) ? __ldg(A + (m0 + m) * K + k0 + k) : 0.f;
      }
      if (B_NK) {
        const int n = e / 16, k = e % 16;
        rb[i] = (n0 + n < N && k0 + k < k_end) ? __ldg(B + (size_t)(n0 + n) * K + k0 + k) : 0.f;
      } else {
        const int k = e / 64, n = e % 64;
        rb[i] = (n0 + n < N && k0 + k < k_end) ? __ldg(B + (size_t)(k0 + k) * N + n0 + n) : 0.f;
      }
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      As[buf][e % 16][e / 16] = ra[i];
      if (B_NK) Bs[buf][e % 16][e / 16] = rb[i];
      else Bs[buf][e / 64][e % 64] = rb[i];
    }
  };
  if (k_begin < k_end) {
    fetch(k_begin);
    stage(0);
  }
  __syncthreads();
  int buf = 0;
  for (int k0 = k_begin; k0 < k_end; k0 += 16, buf ^= 1) {
    const bool more = k0 + 16 < k_end;
    if (more) fetch(k0 + 16);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) stage(buf ^ 1);  // the other buffer was last read before the previous barrier
    __syncthreads();
  }
  const float sc = scale != nullptr ? __ldg(scale) : 1.f;
  const bool add_bias = bias != nullptr && (!SPLITK || blockIdx.z == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const float v = acc[i][j] * sc + (add_bias ? bias[n] : 0.f);
      if (SPLITK) atomicAdd(C + m * N + n, v);
      else C[m * N + n] = v;
    }
  }
}

// Fused-loss finalize: ||r|| = sqrt(sum of the head warps' partial sums of r^2), summed in double in a fixed order.
// norm[0] = ||r||, norm[1] = 1/||r|| (0 when ||r|| == 0: the subgradient torch.linalg.norm's backward uses).
// `extra_sq` (optional, device): a term added to the sum of squares -- the measurement energy of rows that were not
// decoded because their mask weight is zero (r = y_meas there, whatever the network says).
__global__ void __launch_bounds__(1024) loss_finalize_kernel(const float* __restrict__ partials, int n,
                                                             float* __restrict__ norm,
                                                             const float* __restrict__ extra_sq) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)partials[i];
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (threadIdx.x == 0) {
      if (extra_sq != nullptr) v += fmax((double)*extra_sq, 0.0);
      const double nrm = sqrt(v);
      norm[0] = (float)nrm;
      norm[1] = nrm > 0.0 ? (float)(1.0 / nrm) : 0.f;
    }
  }
}

// Fused-loss epilogue of the fp32 (CUDA-core) path as its own pass over the decoded field: residual, seed and partial
// sums exactly as the tensor-core kernels' heads do (tc_common.cuh: tc_loss_row).  One partial slot per block.
__global__ void __launch_bounds__(256) loss_rows_kernel(LossArgs la, const float* __restrict__ y, int64_t T, int64_t P,
                                                        int cout) {
  __shared__ float red[8];
  float sq = 0.f;
  const int64_t rows = T * P;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < rows; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = q / P, p = q - t * P;
    float ys[4] = {0.f, 0.f, 0.f, 0.f};
    for (int o = 0; o < cout; ++o) ys[o] = y[q * cout + o];
    sq += tc_loss_row(la, t, p, P, cout, true, ys);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w];
    la.partials[blockIdx.x % kLossPartials] = v;
  }
}

// ------------------------------------------------------------------------------------------
constexpr int kSimtTM = 64;  // points per block

// out = A(TMxH, smem) * W(HxH, global, [k][n] layout), one 128-column block at a time.
// Thread (ty,tx): rows ty*4..+3, columns cb*128 + tx + 16*j.
template <typename Epi>
__device__ __forceinline__ void simt_layer(const float* __restrict__ hin, int ld, const float* __restrict__ Wkn, int H,
                                           Epi epi) {
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  for (int cb = 0; cb < H; cb += 128) {
    float acc[4][8] = {};
    for (int k = 0; k < H; ++k) {
      float a[4], w[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = hin[(ty * 4 + i) * ld + k];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = cb + tx + 16 * j;
        w[j] = (n < H) ? __ldg(Wkn + (size_t)k * H + n) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = cb + tx + 16 * j;
        if (n < H) epi(ty * 4 + i, n, acc[i][j]);
      }
  }
}

// Forward chain in fp32.  Dynamic smem: 2 * TM * (H+1) floats + TM * cin floats.
template <bool STASH>
__global__ void __launch_bounds__(256) simt_forward_kernel(cnf_dims d, const uint8_t* __restrict__ packed,
                                                           const float* __restrict__ coords,
                                                           int64_t coord_frame_stride, const float* __restrict__ shift,
                                                           float* __restrict__ out, float* __restrict__ stash,
                                                           int64_t T, int64_t P) {
  extern __shared__ float smem_f[];
  const PackedLayout lay = make_layout(d);
  const int H = d.H, nl = d.nl, cin = d.cin, cout = d.cout, ld = H + 1;
  float* hA = smem_f;
  float* hB = hA + kSimtTM * ld;
  float* xs = hB + kSimtTM * ld;
  const float* w_first = reinterpret_cast<const float*>(packed + lay.w_first);
  const float* w_out = reinterpret_cast<const float*>(packed + lay.w_out);
  const float* b_out = reinterpret_cast<const float*>(packed + lay.b_out);
  const float* w_hid_t = reinterpret_cast<const float*>(packed + lay.w_hid_t);
  const int64_t PB = (P + kSimtTM - 1) / kSimtTM;
  const int64_t tiles = T * PB;
  const int tid = threadIdx.x;
  const int64_t SH = (int64_t)(nl + 1) * H;

  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t t = tile / PB, p0 = (tile % PB) * kSimtTM;
    const float* cbase = coords + t * coord_frame_stride;
    const float* sh = shift + t * SH;
    __syncthreads();  // previous tile's readers of xs / h buffers are done
    for (int e = tid; e < kSimtTM * cin; e += 256) {
      const int64_t p = p0 + e / cin;
      xs[e] = (p < P) ? cbase[p * cin + e % cin] : 0.f;
    }
    __syncthreads();
    // layer 0 (K = cin)
    for (int e = tid; e < kSimtTM * H; e += 256) {
      const int r = e / H, n = e % H;
      float z = sh[n];
      for (int j = 0; j < cin; ++j) z = fmaf(w_first[n * cin + j], xs[r * cin + j], z);
      float s, c;
      sincosf(z, &s, &c);
      hA[r * ld + n] = s;
      if (STASH && p0 + r < P) stash[((t * P + p0 + r) * (nl + 1) + 0) * H + n] = c;
    }
    __syncthreads();
    float* hin = hA;
    float* hout = hB;
    for (int l = 1; l <= nl; ++l) {
      const float* Wkn = w_hid_t + (size_t)(l - 1) * H * H;
      const float* shl = sh + (size_t)l * H;
      simt_layer(hin, ld, Wkn, H, [&](int r, int n, float acc) {
        float s, c;
        sincosf(acc + shl[n], &s, &c);
        hout[r * ld + n] = s;
        if (STASH && p0 + r < P) stash[((t * P + p0 + r) * (nl + 1) + l) * H + n] = c;
      });
      __syncthreads();
      float* tmp = hin;
      hin = hout;
      hout = tmp;
    }
    // linear head
    for (int e = tid; e < kSimtTM * cout; e += 256) {
      const int r = e / cout, o = e % cout;
      if (p0 + r >= P) continue;
      float y = b_out[o];
      for (int n = 0; n < H; ++n) y = fmaf(w_out[o * H + n], hin[r * ld + n], y);
      out[(t * P + p0 + r) * cout + o] = y;
    }
  }
}

// Backward to the FiLM shifts in fp32.  gshift must be zero on entry.
// Dynamic smem: 2 * TM * (H+1) floats + TM * cout floats.
__global__ void __launch_bounds__(256) simt_backward_kernel(cnf_dims d, const uint8_t* __restrict__ packed,
                                                            const float* __restrict__ gout,
                                                            const float* __restrict__ stash,
                                                            float* __restrict__ gshift, int64_t T, int64_t P) {
  extern __shared__ float smem_f[];
  const PackedLayout lay = make_layout(d);
  const int H = d.H, nl = d.nl, cout = d.cout, ld = H + 1;
  float* dA = smem_f;
  float* dB = dA + kSimtTM * ld;
  float* gy = dB + kSimtTM * ld;
  const float* w_out = reinterpret_cast<const float*>(packed + lay.w_out);
  const float* w_hid = reinterpret_cast<const float*>(packed + lay.w_hid);
  const int64_t PB = (P + kSimtTM - 1) / kSimtTM;
  const int64_t tiles = T * PB;
  const int tid = threadIdx.x;
  const int64_t SH = (int64_t)(nl + 1) * H;

  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t t = tile / PB, p0 = (tile % PB) * kSimtTM;
    float* gs = gshift + t * SH;
    __syncthreads();
    for (int e = tid; e < kSimtTM * cout; e += 256) {
      const int64_t p = p0 + e / cout;
      gy[e] = (p < P) ? gout[(t * P + p) * cout + e % cout] : 0.f;
    }
    __syncthreads();
    // delta_nl = (gy * Wout) .* cos_nl ; column sums -> gshift[nl]
    for (int n = tid; n < H; n += 256) {
      float colsum = 0.f;
      for (int r = 0; r < kSimtTM; ++r) {
        float g = 0.f;
        for (int o = 0; o < cout; ++o) g = fmaf(gy[r * cout + o], w_out[o * H + n], g);
        const float c = (p0 + r < P) ? stash[((t * P + p0 + r) * (nl + 1) + nl) * H + n] : 0.f;
        const float dl = g * c;
        dA[r * ld + n] = dl;
        colsum += dl;
      }
      atomicAdd(gs + (size_t)nl * H + n, colsum);
    }
    __syncthreads();
    float* din = dA;
    float* dout = dB;
    for (int l = nl; l >= 1; --l) {
      // g[r][k_in] = sum_{n_out} delta_l[r][n_out] * W_l[n_out][k_in]; delta_{l-1} = g .* cos_{l-1}
      const float* Wnk = w_hid + (size_t)(l - 1) * H * H;
      simt_layer(din, ld, Wnk, H, [&](int r, int n, float acc) {
        const float c = (p0 + r < P) ? stash[((t * P + p0 + r) * (nl + 1) + (l - 1)) * H + n] : 0.f;
        dout[r * ld + n] = acc * c;
      });
      __syncthreads();
      for (int n = tid; n < H; n += 256) {
        float colsum = 0.f;
        for (int r = 0; r < kSimtTM; ++r) colsum += dout[r * ld + n];
        atomicAdd(gs + (size_t)(l - 1) * H + n, colsum);
      }
      __syncthreads();
      float* tmp = din;
      din = dout;
      dout = tmp;
    }
  }
}

}  // namespace cnf
