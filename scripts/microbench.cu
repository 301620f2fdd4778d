// Micro-benchmarks that size the CNF forward kernel on B200 (sm_100a): tcgen05.mma issue pace with the A operand in
// shared memory vs tensor memory, tcgen05.ld / tcgen05.st throughput, their interference, and MUFU.SIN throughput.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I confild_b200/csrc scripts/microbench.cu -o scripts/microbench
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "ptx.cuh"
#include "tc2_kernels.cuh"

using namespace cnf;

struct Tail {
  uint64_t bar_mma[2];
  uint32_t tmem_base;
};

// flags: bit0 = run MMAs, bit1 = MMA A operand from TMEM, bit2 = run LDTM loops (warps 0-3), bit3 = also STTM,
//        bit4 = LDTM warps 4-7 too (second slot)
__global__ void __launch_bounds__(320, 1) mb_tensor(int flags, int groups, int mma_per_group, int n_dim,
                                                    unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  Tail* tail = reinterpret_cast<Tail*>(smem + 64 * 1024);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < 16 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&tail->bar_mma[0], 1);
    ptx::mbar_init(&tail->bar_mma[1], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 8) {
    ptx::tmem_alloc(&tail->tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tail->tmem_base;
  const bool do_mma = flags & 1, a_tmem = flags & 2, do_ld = flags & 4, do_st = flags & 8, ld2 = flags & 16;
  const bool alt_d = flags & 32, no_acc = flags & 64;
  unsigned long long t0 = 0, t1 = 0;
  if (warp < 8) {
    if (do_ld && (warp < 4 || ld2)) {
      const int wq = warp % 4, g = warp / 4;
      const uint32_t base = tmem + ((uint32_t)(wq * 32) << 16) + g * 256;
      uint32_t acc = 0;
      __syncwarp();
      t0 = clock64();
      for (int it = 0; it < groups; ++it) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(base + c * 32, v);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc ^= v[j];
          if (do_st) {
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = v[j] + it;
            ptx::tmem_st_32x32b_x16(base + 128 + c * 16, w);
            ptx::tmem_st_32x32b_x16(base + 192 + c * 16, w);
          }
        }
        if (do_st) ptx::tmem_wait_st();
      }
      t1 = clock64();
      if (lane == 0) {
        out[16 + warp * 2] = t1 - t0;
        out[16 + warp * 2 + 1] = acc;
      }
    }
  } else if (warp == 8 && do_mma && (flags & 128)) {
    // warp-uniform issue: all lanes run the loop, one elected lane issues (operands stay in uniform registers)
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = ptx::make_idesc_f16(1u, 128, n_dim);
    const uint64_t a_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem));
    const uint64_t b_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem) + 32 * 1024);
    uint32_t phase[2] = {0u, 0u};
    t0 = clock64();
    for (int it = 0; it < groups; ++it) {
      const uint32_t d = tm + (it & 1) * 256;
      if (ptx::elect_one()) {
        if (flags & 256) {  // two independent N=n_dim accumulators, alternating (12 MMAs each)
#pragma unroll
          for (int k = 0; k < 24; ++k) {
            const uint32_t dd = tm + (k & 1) * 256;
            ptx::umma_f16_ts(dd, tm + 128 + (k & 7) * 8, b_desc + 2 * (k & 3), idesc, k > 1);
          }
        } else if (flags & 512) {  // one N=n_dim accumulator issued as two N/2 column halves, alternating (48 MMAs)
          const uint32_t ih = ptx::make_idesc_f16(1u, 128, n_dim / 2);
          const uint64_t bh = (uint64_t)((n_dim / 2) * 128 / 16);
#pragma unroll
          for (int k = 0; k < 24; ++k) {
            ptx::umma_f16_ts(d, d + 128 + (k & 7) * 8, b_desc + 2 * (k & 3), ih, k != 0);
            ptx::umma_f16_ts(d + n_dim / 2, d + 128 + (k & 7) * 8, b_desc + bh + 2 * (k & 3), ih, k != 0);
          }
        } else if (flags & 1024) {  // same two halves, one after the other (48 MMAs)
          const uint32_t ih = ptx::make_idesc_f16(1u, 128, n_dim / 2);
          const uint64_t bh = (uint64_t)((n_dim / 2) * 128 / 16);
#pragma unroll
          for (int k = 0; k < 24; ++k) ptx::umma_f16_ts(d, d + 128 + (k & 7) * 8, b_desc + 2 * (k & 3), ih, k != 0);
#pragma unroll
          for (int k = 0; k < 24; ++k)
            ptx::umma_f16_ts(d + n_dim / 2, d + 128 + (k & 7) * 8, b_desc + bh + 2 * (k & 3), ih, k != 0);
        } else {
#pragma unroll
          for (int k = 0; k < 24; ++k) {
            if (a_tmem)
              ptx::umma_f16_ts(d, d + 128 + (k & 7) * 8, b_desc + 2 * (k & 3), idesc, k != 0);
            else
              ptx::umma_f16_ss(d, a_desc + 2 * (k & 3), b_desc + 2 * (k & 3), idesc, k != 0);
          }
        }
        ptx::umma_commit(&tail->bar_mma[it & 1]);
      }
      __syncwarp();
      if (it > 0) {
        ptx::mbar_wait(&tail->bar_mma[(it - 1) & 1], phase[(it - 1) & 1]);
        phase[(it - 1) & 1] ^= 1u;
      }
    }
    ptx::mbar_wait(&tail->bar_mma[(groups - 1) & 1], phase[(groups - 1) & 1]);
    t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
  } else if (warp == 8 && lane == 0 && do_mma) {
    const uint32_t idesc = ptx::make_idesc_f16(1u, 128, n_dim);
    const uint64_t a_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem));
    const uint64_t b_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem) + 32 * 1024);
    uint32_t phase[2] = {0u, 0u};
    t0 = clock64();
    for (int it = 0; it < groups; ++it) {
      const uint32_t d0 = tmem + (it & 1) * 256;
      for (int k = 0; k < mma_per_group; ++k) {
        const uint32_t d = alt_d ? tmem + (k & 1) * 256 : d0;
        const uint32_t acc = no_acc ? 0u : (uint32_t)(k > (alt_d ? 1 : 0));
        if (a_tmem)
          ptx::umma_f16_ts(d, d + 128 + (k & 7) * 8, b_desc + 2 * (k & 3), idesc, acc);
        else
          ptx::umma_f16_ss(d, a_desc + 2 * (k & 3), b_desc + 2 * (k & 3), idesc, acc);
      }
      ptx::umma_commit(&tail->bar_mma[it & 1]);
      if (it > 0) {  // keep one group queued behind the one executing
        ptx::mbar_wait(&tail->bar_mma[(it - 1) & 1], phase[(it - 1) & 1]);
        phase[(it - 1) & 1] ^= 1u;
      }
    }
    ptx::mbar_wait(&tail->bar_mma[(groups - 1) & 1], phase[(groups - 1) & 1]);
    t1 = clock64();
    out[0] = t1 - t0;
  }
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// MUFU.SIN throughput: `chains` independent values per thread, `iters` dependent steps each.
template <int CHAINS>
__global__ void mb_mufu(int iters, float seed, float* sink, unsigned long long* cyc) {
  float x[CHAINS];
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) x[j] = seed + 0.01f * (threadIdx.x + 37 * j);
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) x[j] = ptx::sin_approx(x[j]);
  }
  const unsigned long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) s += x[j];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Epilogue alone: `nwarps` warps each run the hidden-layer epilogue of the H=128 kernel back to back
// (no MMA, no hand-shakes), to size E = epilogue time per layer as a function of warps per SM sub-partition.
template <int PREC>
__global__ void __launch_bounds__(576, 1) mb_epilogue(int iters, unsigned long long* out) {
  __shared__ float sbuf[128];
  __shared__ float wout[4 * 128];
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t dummy_bar[2];
  if (threadIdx.x == 0) {
    ptx::mbar_init(&dummy_bar[0], 256);
    ptx::mbar_init(&dummy_bar[1], 256);
    ptx::fence_mbar_init();
  }
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < 128; i += blockDim.x) sbuf[i] = 0.001f * i;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) wout[i] = 0.01f;
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int g = (warp / 8) & 1, hf = (warp / 4) & 1, wq = warp % 4;
  const uint32_t lane_base = tmem + ((uint32_t)(wq * 32) << 16) + g * 256;
  float y[4] = {0.f, 0.f, 0.f, 0.f};
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    tc2_hidden_layer<PREC, false, false, false>(lane_base, lane_base + 128, hf, sbuf, wout, 3, y, nullptr, &dummy_bar[0], &dummy_bar[1]);
    ptx::tmem_wait_st();
  }
  const unsigned long long t1 = clock64();
  if (lane == 0) out[warp] = t1 - t0;
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}


// Sweep: one accumulator of N_TOTAL columns issued as N_TOTAL/N_MMA instructions of N=N_MMA per K step, fully unrolled.
//   ORDER 0: for k { for part }   ORDER 1: for part { for k }.  KSTEPS K steps per group; A from TMEM (TS) or smem.
template <int N_TOTAL, int N_MMA, int ORDER, int KSTEPS, int TS, int BG = 0>
__global__ void __launch_bounds__(320, 1) mb_sweep(int groups, unsigned long long* out) {
  __shared__ float sbuf[128];
  __shared__ float wout[4 * 128];
  __shared__ uint64_t dummy_bar[2];
  if (threadIdx.x == 0) {
    ptx::mbar_init(&dummy_bar[0], 256);
    ptx::mbar_init(&dummy_bar[1], 256);
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) sbuf[i] = 0.001f * i;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) wout[i] = 0.01f;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  Tail* tail = reinterpret_cast<Tail*>(smem + 96 * 1024);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < 24 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&tail->bar_mma[0], 1);
    ptx::mbar_init(&tail->bar_mma[1], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 8) {
    ptx::tmem_alloc(&tail->tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 8) {
    const uint32_t tm = __shfl_sync(0xffffffffu, tail->tmem_base, 0);
    const uint32_t idesc = ptx::make_idesc_f16(1u, 128, N_MMA);
    const uint64_t a_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem));
    const uint64_t b_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem) + 32 * 1024);
    constexpr int kParts = N_TOTAL / N_MMA;
    constexpr uint32_t kBStep = N_MMA * 128 / 16;
    const uint32_t a_tm = BG ? tm + 128 : tm + 384;
    uint32_t phase[2] = {0u, 0u};
    const unsigned long long t0 = clock64();
    for (int it = 0; it < groups; ++it) {
      if (ptx::elect_one()) {
#pragma unroll
        for (int i = 0; i < kParts * KSTEPS; ++i) {
          const int k = ORDER ? i % KSTEPS : i / kParts;
          const int part = ORDER ? i / KSTEPS : i % kParts;
          if (TS)
            ptx::umma_f16_ts(tm + part * N_MMA, a_tm + (k & 7) * 8, b_desc + part * kBStep + 2 * (k & 3), idesc, k != 0);
          else
            ptx::umma_f16_ss(tm + part * N_MMA, a_desc + 2 * (k & 3), b_desc + part * kBStep + 2 * (k & 3), idesc, k != 0);
        }
        ptx::umma_commit(&tail->bar_mma[it & 1]);
      }
      __syncwarp();
      if (it > 0) {
        ptx::mbar_wait(&tail->bar_mma[(it - 1) & 1], phase[(it - 1) & 1]);
        phase[(it - 1) & 1] ^= 1u;
      }
    }
    ptx::mbar_wait(&tail->bar_mma[(groups - 1) & 1], phase[(groups - 1) & 1]);
    const unsigned long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
  } else if (BG) {
    // background: the real H=128 activation epilogue of the OTHER tile slot (TMEM columns 256..511), 8 warps
    const int hf = (warp / 4) & 1, wq = warp % 4;
    const uint32_t lane_base = tail->tmem_base + ((uint32_t)(wq * 32) << 16) + 256;
    float y[4] = {0.f, 0.f, 0.f, 0.f};
    const unsigned long long t0 = clock64();
    const int iters = groups * 2;
    for (int it = 0; it < iters; ++it) {
      if (BG == 1) tc2_hidden_layer<CNF_PREC_BF16X3, false, false, false>(lane_base, lane_base + 128, hf, sbuf, wout, 3, y, nullptr, &dummy_bar[0], &dummy_bar[1]);
      else tc2_hidden_layer<CNF_PREC_FP16, false, false, false>(lane_base, lane_base + 128, hf, sbuf, wout, 3, y, nullptr, &dummy_bar[0], &dummy_bar[1]);
      ptx::tmem_wait_st();
    }
    const unsigned long long t1 = clock64();
    if (lane == 0) out[8 + warp] = (t1 - t0) / iters;
    if (y[0] == 123.456f) out[40] = 1;
  }
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tail->tmem_base, 512);
  }
}

template <int N_TOTAL, int N_MMA, int ORDER, int KSTEPS, int TS, int BG = 0>
void run_sweep(unsigned long long* d_out) {
  const size_t sm2 = 96 * 1024 + 1024 + sizeof(Tail);
  cudaFuncSetAttribute(mb_sweep<N_TOTAL, N_MMA, ORDER, KSTEPS, TS, BG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
  const int groups = 64;
  unsigned long long h = 0;
  cudaMemset(d_out, 0, 64 * sizeof(unsigned long long));
  mb_sweep<N_TOTAL, N_MMA, ORDER, KSTEPS, TS, BG><<<148, 320, sm2>>>(groups, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
  if (BG) {
    unsigned long long e = 0;
    cudaMemcpy(&e, d_out + 8, 8, cudaMemcpyDeviceToHost);
    printf("[background epilogue %s, 8 warps: %llu clk per layer] ", BG == 1 ? "bf16x3" : "fp16", e);
  }
  printf("%s n_total=%3d ksteps=%2d n_mma=%3d order=%s : %7.1f clk per K step (%6.1f per MMA)\n", TS ? "TS" : "SS", N_TOTAL,
         KSTEPS, N_MMA, ORDER ? "part-major" : "k-major   ", (double)h / (groups * KSTEPS),
         (double)h / (groups * KSTEPS * (N_TOTAL / N_MMA)));
}

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

int main(int argc, char** argv) {
  unsigned long long* d_out;
  CK(cudaMalloc(&d_out, 64 * sizeof(unsigned long long)));
  const size_t smem = 64 * 1024 + 1024 + sizeof(Tail);
  CK(cudaFuncSetAttribute(mb_tensor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<unsigned long long> h(64);
  if (argc > 1 && std::string(argv[1]) == "sweep") {
#define SW(NT, NM, TS) run_sweep<NT, NM, 0, 24, TS>(d_out); if (NT != NM) run_sweep<NT, NM, 1, 24, TS>(d_out);
    SW(128, 32, 1) SW(128, 64, 1) SW(128, 128, 1)
    SW(256, 64, 1) SW(256, 128, 1) SW(256, 256, 1)
    SW(128, 32, 0) SW(128, 64, 0) SW(128, 128, 0)
    SW(256, 64, 0) SW(256, 128, 0) SW(256, 256, 0)
    SW(384, 64, 0) SW(384, 96, 0) SW(384, 128, 0) SW(384, 192, 0)
    run_sweep<128, 64, 0, 48, 1>(d_out); run_sweep<128, 128, 0, 48, 1>(d_out);
    run_sweep<128, 128, 0, 24, 1, 1>(d_out); run_sweep<128, 64, 0, 24, 1, 1>(d_out); run_sweep<128, 128, 0, 24, 1, 2>(d_out);
    run_sweep<128, 128, 0, 24, 0, 1>(d_out);
    return 0;
  }
  struct Cfg { const char* name; int flags; int n; };
  const Cfg cfgs[] = {
      {"MMA SS  N=128 alone", 1, 128},       {"MMA TS  N=128 alone", 3, 128},
      {"MMA SS  N=256 alone", 1, 256},       {"MMA TS  N=256 alone", 3, 256},
      {"LDTM 4 warps alone", 4, 128},        {"LDTM+STTM 4 warps alone", 12, 128},
      {"LDTM 8 warps alone", 20, 128},       {"LDTM+STTM 8 warps alone", 28, 128},
      {"MMA SS N=128 + LDTM 4w", 5, 128},    {"MMA TS N=128 + LDTM 4w", 7, 128},
      {"MMA SS N=128 + LDTM+STTM 4w", 13, 128}, {"MMA TS N=128 + LDTM+STTM 4w", 15, 128},
      {"MMA TS N=128 + LDTM+STTM 8w", 31, 128},
      {"MMA SS N=128 uniform issue", 1 | 128, 128}, {"MMA TS N=128 uniform issue", 3 | 128, 128},
      {"MMA TS N=128 uniform + LDTM+STTM 8w", 31 | 128, 128}, {"MMA SS N=128 uniform + LDTM+STTM 8w", 29 | 128, 128},
      {"MMA TS N=256 uniform issue", 3 | 128, 256}, {"MMA SS N=256 uniform issue", 1 | 128, 256},
      {"MMA TS N=64 uniform issue", 3 | 128, 64}, {"MMA SS N=64 uniform issue", 1 | 128, 64},
      {"MMA TS N=32 uniform issue", 3 | 128, 32}, {"MMA TS N=192 uniform issue", 3 | 128, 192},
      {"MMA TS N=128 uniform, 2 accumulators alternating", 3 | 128 | 256, 128},
      {"MMA TS N=128 uniform as 2x N=64 halves alternating (per 24-MMA-equivalent)", 3 | 128 | 512, 128},
      {"MMA TS N=128 uniform as 2x N=64 halves sequential (per 24-MMA-equivalent)", 3 | 128 | 1024, 128},
      {"MMA TS N=256 uniform as 2x N=128 halves alternating (per 24-MMA-equivalent)", 3 | 128 | 512, 256},
      {"MMA SS N=128 alternate D", 1 | 32, 128}, {"MMA TS N=128 alternate D", 3 | 32, 128},
      {"MMA SS N=128 no accumulate", 1 | 64, 128}, {"MMA TS N=128 no accumulate", 3 | 64, 128},
      {"MMA SS N=64", 1, 64}, {"MMA TS N=64", 3, 64}, {"MMA SS N=32", 1, 32}, {"MMA SS N=16", 1, 16},
      {"MMA SS N=192", 1, 192}, {"MMA TS N=64 alternate D", 3 | 32, 64},
  };
  const int groups = 64, per_group = 24;
  for (const Cfg& c : cfgs) {
    CK(cudaMemset(d_out, 0, 64 * sizeof(unsigned long long)));
    mb_tensor<<<1, 320, smem>>>(c.flags, groups, per_group, c.n, d_out);
    CK(cudaDeviceSynchronize());
    mb_tensor<<<148, 320, smem>>>(c.flags, groups, per_group, c.n, d_out);  // whole chip, CTA 0..147 overwrite the same slots
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d_out, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    printf("%-32s", c.name);
    if (c.flags & 1) printf(" mma: %7.1f clk/MMA (%llu clk / %d)", (double)h[0] / (groups * per_group), h[0], groups * per_group);
    if (c.flags & 4) {
      const double clk = (double)h[16];
      const double bytes = (double)groups * 4 * 4096;  // per warp: 4 x LDTM.x32 of 4 KiB per group
      printf("  ldtm warp0: %7.1f clk/LDTM.x32 (%.1f B/clk/warp, x%d warps)", clk / (groups * 4), bytes / clk,
             (c.flags & 16) ? 8 : 4);
    }
    printf("\n");
  }
  for (int warps : {4, 8, 16}) {
    const int iters = 200;
    CK(cudaMemset(d_out, 0, 64 * sizeof(unsigned long long)));
    mb_epilogue<CNF_PREC_BF16X3><<<148, warps * 32>>>(iters, d_out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d_out, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    printf("epilogue bf16x3 %2d warps/SM: warp0 %7.1f clk per 64-column layer epilogue (last warp %7.1f)\n", warps,
           (double)h[0] / iters, (double)h[warps - 1] / iters);
    mb_epilogue<CNF_PREC_FP16><<<148, warps * 32>>>(iters, d_out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d_out, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    printf("epilogue fp16   %2d warps/SM: warp0 %7.1f clk per 64-column layer epilogue (last warp %7.1f)\n", warps,
           (double)h[0] / iters, (double)h[warps - 1] / iters);
  }
  // MUFU
  float* d_sink;
  CK(cudaMalloc(&d_sink, 148 * 1024 * sizeof(float)));
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 2000;
    mb_mufu<8><<<148, warps * 32>>>(iters, 0.3f, d_sink, d_out);
    CK(cudaDeviceSynchronize());
    mb_mufu<8><<<148, warps * 32>>>(iters, 0.3f, d_sink, d_out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d_out, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    const double sins = (double)iters * 8 * warps * 32;
    printf("MUFU.SIN %2d warps/SM, 8 chains: %8llu clk -> %.2f sin/clk/SM\n", warps, h[0], sins / (double)h[0]);
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int iters = 20000;
  mb_mufu<8><<<148 * 2, 512>>>(iters, 0.3f, d_sink, d_out);
  CK(cudaEventRecord(e0));
  mb_mufu<8><<<148 * 2, 512>>>(iters, 0.3f, d_sink, d_out);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("MUFU.SIN chip-wide: %.3f T sin/s (%.3f ms for %.3e sins)\n", (double)iters * 8 * 512 * 296 / (ms * 1e-3) / 1e12, ms,
         (double)iters * 8 * 512 * 296);
  return 0;
}
