// Sampler-side helper: GroupNorm (+ optional per-(sample, channel) pre-add, + optional SiLU) on
// channels-last bf16 activations, statistics in fp32.
//
// Replaces, per normalisation site of the guided-diffusion U-Net's residual / attention blocks
// (UnconditionalDiffusionTraining_and_Generation/src/unet.py:185-200,228-256,283-300 with src/nn.py:17-19 GroupNorm32),
// the eager chain  x.float() -> native_group_norm (moments, fused params, apply; NCHW only) -> .to(bf16) -> SiLU
// and the NCHW <-> NHWC conversions cuDNN then runs around every convolution: 6-8 launches and ~5 passes over the
// activation become one read for the statistics and one read + one write for the result.
//
// Layout: x, y = [N][HW][C] bf16 (a torch channels_last (N,C,H,W) tensor); thread (tx, ty) of a block owns the 16-byte
// vector of channels [8 tx, 8 tx + 8) of every R-th pixel of the block's pixel range (R = blockDim.y), so every warp
// load / store is a fully coalesced run of the tensor.  Two kernels, no atomics (deterministic):
//   gn_stats  per (chunk of pixels, sample): per-channel sums -> shared memory -> per-group (sum, sum of squares)
//             written to partials[n][chunk][g][2]
//   gn_apply  every block re-reduces its sample's <= 128 chunk partials into (mean, rstd), folds gamma / beta / the
//             pre-add into one (a_c, b_c) pair per channel and streams y = act(a_c x + b_c).
// Small activations (<= 256 KiB per sample: the 8x8 / 16x16 levels of the U-Net) take ONE kernel instead: gn_cluster, one
// thread-block cluster of 8 CTAs per sample -- statistics of the CTA's pixels -> shared memory, barrier.cluster, every
// CTA sums the 8 partial results over distributed shared memory in a fixed order, applies.  Measured over the 71 sites
// of the case1 U-Net replayed from a CUDA graph (tests/tools/gn_ab.py): 864 us two kernels everywhere, 811 us with the
// cluster path up to 256 KiB, 887 / 1,017 / 1,102 us with it up to 512 KiB / 1 MiB / 2 MiB (16 CTAs are too few there).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "host.cuh"
#include "ptx.cuh"

namespace cnf {
constexpr int kGnMaxChunks = CNF_GN_MAX_CHUNKS;
constexpr int kGnMaxGroups = 64;
namespace {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

__global__ void __launch_bounds__(256) gn_stats_kernel(const uint4* __restrict__ x, const float* __restrict__ add,
                                                       int64_t add_stride, float* __restrict__ partials, int HW, int C,
                                                       int G, int chunk_px, int chunks) {
  extern __shared__ __align__(16) float sm[];  // [2][R][C]
  const int vecs = C / 8, R = blockDim.y, tx = threadIdx.x, ty = threadIdx.y;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int p0 = chunk * chunk_px, p1 = min(HW, p0 + chunk_px);
  float e[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (add != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] = add[(size_t)n * add_stride + tx * 8 + j];
  }
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const uint4* xp = x + (size_t)n * HW * vecs;
  for (int p = p0 + ty; p < p1; p += R) {
    float f[8];
    unpack8(__ldg(xp + (size_t)p * vecs + tx), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = f[j] + e[j];
      s[j] += v;
      q[j] = fmaf(v, v, q[j]);
    }
  }
  {  // 16-byte stores: conflict-free per quarter-warp
    float4* ps = reinterpret_cast<float4*>(sm + (size_t)ty * C + tx * 8);
    float4* pq = reinterpret_cast<float4*>(sm + (size_t)(R + ty) * C + tx * 8);
    ps[0] = make_float4(s[0], s[1], s[2], s[3]);
    ps[1] = make_float4(s[4], s[5], s[6], s[7]);
    pq[0] = make_float4(q[0], q[1], q[2], q[3]);
    pq[1] = make_float4(q[4], q[5], q[6], q[7]);
  }
  __syncthreads();
  const int tid = ty * vecs + tx;
  if (tid < G) {
    const int Cg = C / G;
    float S = 0.f, Q = 0.f;
    for (int r = 0; r < R; ++r)
      for (int c = tid * Cg; c < (tid + 1) * Cg; ++c) {
        S += sm[(size_t)r * C + c];
        Q += sm[(size_t)(R + r) * C + c];
      }
    float* out = partials + (((size_t)n * chunks + chunk) * G + tid) * 2;
    out[0] = S;
    out[1] = Q;
  }
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const uint4* __restrict__ x, const float* __restrict__ add,
                                                       int64_t add_stride, const float* __restrict__ partials,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       uint4* __restrict__ y, int HW, int C, int G, int chunk_px,
                                                       int chunks, float eps, int silu) {
  extern __shared__ __align__(16) float sm[];  // a[C], b[C], mean[G], rstd[G]
  float* a = sm;
  float* b = sm + C;
  float* mean = b + C;
  float* rstd = mean + G;
  const int vecs = C / 8, R = blockDim.y, tx = threadIdx.x, ty = threadIdx.y;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int tid = ty * vecs + tx, nthreads = vecs * R;
  if (tid < G) {
    float S = 0.f, Q = 0.f;
    for (int c = 0; c < chunks; ++c) {  // fixed order: the same (mean, rstd) in every block, run to run
      const float* in = partials + (((size_t)n * chunks + c) * G + tid) * 2;
      S += in[0];
      Q += in[1];
    }
    const float inv_cnt = 1.f / ((float)HW * (float)(C / G));
    const float m = S * inv_cnt;
    const float var = fmaxf(fmaf(-m, m, Q * inv_cnt), 0.f);
    mean[tid] = m;
    rstd[tid] = rsqrtf(var + eps);
  }
  __syncthreads();
  for (int c = tid; c < C; c += nthreads) {
    const int g = c / (C / G);
    const float ga = gamma[c] * rstd[g];
    a[c] = ga;
    b[c] = fmaf((add != nullptr ? add[(size_t)n * add_stride + c] : 0.f) - mean[g], ga, beta[c]);
  }
  __syncthreads();
  float ra[8], rb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ra[j] = a[tx * 8 + j];
    rb[j] = b[tx * 8 + j];
  }
  const int p0 = chunk * chunk_px, p1 = min(HW, p0 + chunk_px);
  const uint4* xp = x + (size_t)n * HW * vecs;
  uint4* yp = y + (size_t)n * HW * vecs;
  for (int p = p0 + ty; p < p1; p += R) {
    float f[8];
    unpack8(__ldg(xp + (size_t)p * vecs + tx), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = fmaf(f[j], ra[j], rb[j]);
      if (silu) v = __fdividef(v, 1.f + __expf(-v));
      f[j] = v;
    }
    uint4 o;
    __nv_bfloat162 h;
    h = __floats2bfloat162_rn(f[0], f[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(f[2], f[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(f[4], f[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(f[6], f[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
    yp[(size_t)p * vecs + tx] = o;
  }
}

constexpr int kGnCluster = 8;

__global__ void __launch_bounds__(256) gn_cluster_kernel(const uint4* __restrict__ x, const float* __restrict__ add,
                                                         int64_t add_stride, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, uint4* __restrict__ y, int HW,
                                                         int C, int G, int chunk_px, float eps, int silu) {
  extern __shared__ __align__(16) float sm[];  // [2][R][C] channel sums, reused as a[C], b[C]; then part[G][2], stat[G][2]
  const int vecs = C / 8, R = blockDim.y, tx = threadIdx.x, ty = threadIdx.y;
  const int n = blockIdx.y;
  const uint32_t rank = ptx::cluster_ctarank();
  float* part = sm + (size_t)2 * R * C;  // this CTA's per-group (sum, sum of squares): read by the whole cluster
  float* stat = part + 2 * G;            // (mean, rstd) per group
  const int p0 = (int)rank * chunk_px, p1 = min(HW, p0 + chunk_px);
  const int tid = ty * vecs + tx, nthreads = vecs * R;
  float e[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (add != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] = add[(size_t)n * add_stride + tx * 8 + j];
  }
  const uint4* xp = x + (size_t)n * HW * vecs;
  {
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p = p0 + ty; p < p1; p += R) {
      float f[8];
      unpack8(__ldg(xp + (size_t)p * vecs + tx), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = f[j] + e[j];
        s[j] += v;
        q[j] = fmaf(v, v, q[j]);
      }
    }
    float4* ps = reinterpret_cast<float4*>(sm + (size_t)ty * C + tx * 8);
    float4* pq = reinterpret_cast<float4*>(sm + (size_t)(R + ty) * C + tx * 8);
    ps[0] = make_float4(s[0], s[1], s[2], s[3]);
    ps[1] = make_float4(s[4], s[5], s[6], s[7]);
    pq[0] = make_float4(q[0], q[1], q[2], q[3]);
    pq[1] = make_float4(q[4], q[5], q[6], q[7]);
  }
  __syncthreads();
  if (tid < G) {
    const int Cg = C / G;
    float S = 0.f, Q = 0.f;
    for (int r = 0; r < R; ++r)
      for (int c = tid * Cg; c < (tid + 1) * Cg; ++c) {
        S += sm[(size_t)r * C + c];
        Q += sm[(size_t)(R + r) * C + c];
      }
    part[2 * tid] = S;
    part[2 * tid + 1] = Q;
  }
  ptx::cluster_sync_all();  // every CTA's part[] is written (release / acquire at cluster scope)
  if (tid < G) {
    float S = 0.f, Q = 0.f;
    const uint32_t local = ptx::smem_u32(part + 2 * tid);
    for (uint32_t r = 0; r < (uint32_t)kGnCluster; ++r) {  // fixed order: identical (mean, rstd) in all CTAs, run to run
      const uint32_t remote = ptx::mapa_shared(local, r);
      float a, b;
      asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(remote));
      S += a;
      Q += b;
    }
    const float inv_cnt = 1.f / ((float)HW * (float)(C / G));
    const float m = S * inv_cnt;
    stat[2 * tid] = m;
    stat[2 * tid + 1] = rsqrtf(fmaxf(fmaf(-m, m, Q * inv_cnt), 0.f) + eps);
  }
  __syncthreads();
  float* a = sm;  // the channel sums are consumed: reuse
  float* b = sm + C;
  for (int c = tid; c < C; c += nthreads) {
    const int g = c / (C / G);
    const float ga = gamma[c] * stat[2 * g + 1];
    a[c] = ga;
    b[c] = fmaf((add != nullptr ? add[(size_t)n * add_stride + c] : 0.f) - stat[2 * g], ga, beta[c]);
  }
  __syncthreads();
  float ra[8], rb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ra[j] = a[tx * 8 + j];
    rb[j] = b[tx * 8 + j];
  }
  uint4* yp = y + (size_t)n * HW * vecs;
  for (int p = p0 + ty; p < p1; p += R) {
    float f[8];
    unpack8(__ldg(xp + (size_t)p * vecs + tx), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = fmaf(f[j], ra[j], rb[j]);
      if (silu) v = __fdividef(v, 1.f + __expf(-v));
      f[j] = v;
    }
    uint4 o;
    __nv_bfloat162 h;
    h = __floats2bfloat162_rn(f[0], f[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(f[2], f[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(f[4], f[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(f[6], f[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
    yp[(size_t)p * vecs + tx] = o;
  }
  ptx::cluster_sync_all();  // nobody's part[] may disappear (CTA exit) while a peer still reads it
}

}  // namespace
}  // namespace cnf

extern "C" size_t cnf_group_norm_scratch_bytes(int64_t N) {
  return N <= 0 ? 0 : (size_t)N * cnf::kGnMaxChunks * cnf::kGnMaxGroups * 2 * sizeof(float);
}

extern "C" int cnf_group_norm_nhwc_bf16(const void* d_x, const float* d_add, int64_t add_stride, const float* d_gamma,
                                        const float* d_beta, void* d_y, float* d_partials, int64_t N, int64_t HW,
                                        int32_t C, int32_t groups, float eps, int32_t silu, void* stream) {
  using namespace cnf;
  using cnf::host::fail;
  if (d_x == nullptr || d_y == nullptr || d_gamma == nullptr || d_beta == nullptr || d_partials == nullptr)
    return fail(CNF_ERR_INVALID_ARGUMENT, "cnf_group_norm_nhwc_bf16: null pointer");
  if (N <= 0 || HW <= 0) return CNF_OK;
  if (C <= 0 || C % 8 != 0 || C > 2048 || groups <= 0 || groups > kGnMaxGroups || C % groups != 0)
    return fail(CNF_ERR_UNSUPPORTED, "cnf_group_norm_nhwc_bf16: C = %d must be a multiple of 8 (<= 2048) and of groups = %d (<= %d)",
                (int)C, (int)groups, kGnMaxGroups);
  if (N > 65535 || HW > (int64_t)1 << 30) return fail(CNF_ERR_UNSUPPORTED, "cnf_group_norm_nhwc_bf16: N or HW too large");
  if ((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_y)) & 15)
    return fail(CNF_ERR_INVALID_ARGUMENT, "cnf_group_norm_nhwc_bf16: x and y must be 16-byte aligned");
  const int vecs = C / 8;
  int R = 256 / vecs;
  if (R < 1) R = 1;
  // enough pixel chunks to fill the GPU, at least ~4 pixels per thread row, at most kGnMaxChunks (the scratch layout)
  int64_t chunks = (HW + (int64_t)R * 4 - 1) / ((int64_t)R * 4);
  if (chunks > kGnMaxChunks) chunks = kGnMaxChunks;
  if (chunks < 1) chunks = 1;
  const int chunk_px = (int)((HW + chunks - 1) / chunks);
  chunks = (HW + chunk_px - 1) / chunk_px;
  const dim3 grid((unsigned)chunks, (unsigned)N), block((unsigned)vecs, (unsigned)R);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d_add != nullptr && add_stride < C) return fail(CNF_ERR_INVALID_ARGUMENT, "cnf_group_norm_nhwc_bf16: add_stride < C");
  // small activation: one cluster of kGnCluster CTAs per sample, a single launch (see the header comment)
  if ((size_t)HW * C * 2 <= ((size_t)cnf::host::knobs().gn_cluster << 10)) {  // knob = threshold in KiB per sample
    const int cpx = (int)((HW + kGnCluster - 1) / kGnCluster);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kGnCluster, (unsigned)N);
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = ((size_t)2 * R * C + 4 * groups) * sizeof(float);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kGnCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CNF_CUDA(cudaLaunchKernelEx(&cfg, gn_cluster_kernel, static_cast<const uint4*>(d_x), d_add, add_stride, d_gamma, d_beta,
                                static_cast<uint4*>(d_y), (int)HW, (int)C, (int)groups, cpx, eps, (int)silu));
    return CNF_OK;
  }
  const size_t smem_stats = (size_t)2 * R * C * sizeof(float);
  const size_t smem_apply = ((size_t)2 * C + 2 * groups) * sizeof(float);
  gn_stats_kernel<<<grid, block, smem_stats, st>>>(static_cast<const uint4*>(d_x), d_add, add_stride, d_partials, (int)HW,
                                                   C, groups, chunk_px, (int)chunks);
  CNF_CUDA(cudaGetLastError());
  gn_apply_kernel<<<grid, block, smem_apply, st>>>(static_cast<const uint4*>(d_x), d_add, add_stride, d_partials, d_gamma, d_beta,
                                                   static_cast<uint4*>(d_y), (int)HW, C, groups, chunk_px, (int)chunks, eps,
                                                   silu);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}
