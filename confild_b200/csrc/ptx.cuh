// Thin inline-PTX wrappers for the sm_100a features the CNF kernels use:
// mbarrier, 1-D bulk TMA copies (cp.async.bulk), tcgen05 (alloc / mma / commit / ld / fences).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace cnf {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"  // %3: suspend-time hint, the probe sleeps in HW
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680)
      : "memory");
  return ok != 0;
}
// Spin until the phase with the given parity has completed.  A protocol bug would otherwise hang
// the GPU until the watchdog: after 2^22 failed probes trap instead, so the launch fails.
// Same probe with acquire semantics at CLUSTER scope: the arrivals being waited for came from another CTA of the cluster.
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
template <bool CLUSTER_SCOPE = false>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!(CLUSTER_SCOPE ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity))) {
    if ((++spins & 0xFFu) == 0) {  // every 256 failed probes: wall-clock check (a probe may sleep in hardware)
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {  // 4 s without progress: a protocol bug, fail the launch instead of hanging
        printf("cnf: mbarrier wait timed out (block %d thread %d, barrier at shared 0x%x, parity %u)\n", (int)blockIdx.x,
               (int)threadIdx.x, smem_u32(bar), parity);
        while (global_timer_ns() - now < 1000000000ull) {}  // let the other stuck waiters report before the launch dies
        __trap();
      }
    }
  }
}

// One lane of the (converged) warp is elected; the same lane every time.  Keeping the surrounding control flow
// warp-uniform lets the compiler hold tcgen05 operands in uniform registers (no per-thread waterfall loop).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- proxies / TMA
// Make generic-proxy shared-memory writes visible to the async proxy (UMMA operand reads, TMA).
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 1-D bulk copy global -> shared, completion reported as transaction bytes on `bar`.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Same copy delivered to the same CTA-relative shared-memory offset (and mbarrier) of every CTA in `cta_mask` of the
// cluster: one L2 read feeds several SMs.
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                                   uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

// ---------------------------------------------------------------- thread-block clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// Barrier over every thread of every CTA of the cluster (release / acquire): executed by ALL threads, converged.
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Address of `local` (a shared::cta address of THIS CTA) in the shared memory of CTA `rank` of the cluster.
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
// Arrive (release at cluster scope) on an mbarrier given by its shared::cluster address (own or peer CTA).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// ---------------------------------------------------------------- named barriers
__device__ __forceinline__ void bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers fp16 and bf16 operands, fp32 accumulate.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand read from tensor memory (row i of A = TMEM lane i, two 16-bit K elements per column).
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f8f6f4 with 8-bit operands (e4m3 / e5m2 selected per operand in the instruction descriptor): K = 32 per
// instruction, i.e. twice the K of kind::f16 in the same tensor-pipe time.  A in tensor memory: row i = TMEM lane i,
// four 8-bit K elements per 32-bit column (element k in byte k % 4).
__device__ __forceinline__ void umma_f8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f8_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on `bar` when every tcgen05 op issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Same arrival delivered to the mbarrier at the same CTA-relative offset in every CTA of `cta_mask`.
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ---- CTA-pair (cta_group::2) forms: one instruction, issued by a thread of the leader CTA, drives the tensor cores of
// both CTAs of a 2-CTA cluster: M = 256 = 128 rows per CTA (each CTA's own A operand and accumulator), the B operand
// split along N (each CTA's shared memory holds N/2 of its rows).  All addresses are CTA-relative (same in both CTAs).
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {  // one full warp in EACH CTA
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
#define CNF_DEFINE_UMMA_PAIR(name, kind, aop, atype, aconstraint)                                              \
  __device__ __forceinline__ void name(uint32_t tmem_d, atype a, uint64_t desc_b, uint32_t idesc,             \
                                       uint32_t accumulate) {                                                 \
    asm volatile(                                                                                             \
        "{\n\t.reg .pred p;\n\t"                                                                             \
        "setp.ne.b32 p, %4, 0;\n\t"                                                                           \
        "tcgen05.mma.cta_group::2.kind::" kind " [%0], " aop ", %2, %3, p;\n\t}" ::"r"(tmem_d),               \
        aconstraint(a), "l"(desc_b), "r"(idesc), "r"(accumulate)                                              \
        : "memory");                                                                                          \
  }
CNF_DEFINE_UMMA_PAIR(umma_f16_ss_pair, "f16", "%1", uint64_t, "l")
CNF_DEFINE_UMMA_PAIR(umma_f16_ts_pair, "f16", "[%1]", uint32_t, "r")
CNF_DEFINE_UMMA_PAIR(umma_f8_ss_pair, "f8f6f4", "%1", uint64_t, "l")
CNF_DEFINE_UMMA_PAIR(umma_f8_ts_pair, "f8f6f4", "[%1]", uint32_t, "r")
#undef CNF_DEFINE_UMMA_PAIR
// Completion of every cta_group::2 MMA issued so far by this thread, delivered to the mbarrier at the same CTA-relative
// offset in every CTA of `cta_mask`.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive 32-bit columns (thread i gets lane i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 16-column variants (software-pipelined epilogue works on 16 activations at a time)
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns (thread i writes lane i).
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B: rows of 128 bytes (64 x 16-bit),
// 8-row groups 1024 bytes apart (SBO), descriptor version 1 (sm_100), base must be 1024-aligned;
// stepping along K inside the 128-byte row is done by adding the byte offset to the start address.
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= static_cast<uint64_t>(0) << 16;                      // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // stride byte offset, bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                      // descriptor version = 1
  d |= static_cast<uint64_t>(2) << 61;                      // layout type: SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: fp32 accumulate, A and B K-major.
//   fmt: 0 = fp16 operands, 1 = bf16 operands.
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t fmt, uint32_t M, uint32_t N) {
  return (1u << 4)            // D format: fp32
         | (fmt << 7)         // A format
         | (fmt << 10)        // B format
         | (0u << 15)         // A K-major
         | (0u << 16)         // B K-major
         | ((N >> 3) << 17)   // N / 8
         | ((M >> 4) << 24);  // M / 16
}

// Instruction descriptor for kind::f8f6f4: fp32 accumulate, A and B K-major; operand formats 0 = e4m3, 1 = e5m2.
__host__ __device__ constexpr uint32_t make_idesc_f8(uint32_t a_fmt, uint32_t b_fmt, uint32_t M, uint32_t N) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
constexpr uint32_t kF8E4M3 = 0u, kF8E5M2 = 1u;

// ---------------------------------------------------------------- small math helpers
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {  // low half <- lo
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16lo_to_f32(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t packed) { return __uint_as_float(packed & 0xFFFF0000u); }

// Residuals x - bf16(x) of the two values packed in `hi` (cvt.rn.bf16x2: x0 in the low half).
// sm_100 mixed-precision FMA (fma.rn.f32.bf16: 16-bit a, b, fp32 c and result): SASS FHFMA reads the 16-bit half straight
// out of the packed register (.H0 / .H1 selectors), so x - h = fma(h, -1, x) is ONE instruction per value with no unpack
// (before: shift / mask to widen each half + one fma.f32x2 per pair = 1.5 instructions per value).  Exact, as before.
__device__ __forceinline__ float2 bf16x2_residual(uint32_t hi, float x0, float x1) {
  float2 r;
  const unsigned short h0 = (unsigned short)(hi & 0xffffu), h1 = (unsigned short)(hi >> 16);
  asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(r.x) : "h"(h0), "h"((unsigned short)0xBF80), "f"(x0));  // 0xBF80 = -1.0
  asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(r.y) : "h"(h1), "h"((unsigned short)0xBF80), "f"(x1));
  return r;
}

// Four fp32 values -> four 8-bit floats in one word (x0 in byte 0).  cvt.*x2 puts its FIRST source in the upper byte.
__device__ __forceinline__ uint32_t pack_e4m3x4(float x0, float x1, float x2, float x3) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .b16 lo, hi;\n\t"
      "cvt.rn.satfinite.e4m3x2.f32 lo, %2, %1;\n\t"
      "cvt.rn.satfinite.e4m3x2.f32 hi, %4, %3;\n\t"
      "mov.b32 %0, {lo, hi};\n\t}"
      : "=r"(r)
      : "f"(x0), "f"(x1), "f"(x2), "f"(x3));
  return r;
}
__device__ __forceinline__ uint32_t pack_e5m2x4(float x0, float x1, float x2, float x3) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .b16 lo, hi;\n\t"
      "cvt.rn.satfinite.e5m2x2.f32 lo, %2, %1;\n\t"
      "cvt.rn.satfinite.e5m2x2.f32 hi, %4, %3;\n\t"
      "mov.b32 %0, {lo, hi};\n\t}"
      : "=r"(r)
      : "f"(x0), "f"(x1), "f"(x2), "f"(x3));
  return r;
}
// Residuals x - fp16(x) of the two values packed in `hi` (cvt.rn.f16x2: x0 in the low half): one fma.f32x2.
__device__ __forceinline__ float2 f16x2_residual(uint32_t hi, float x0, float x1) {
  float2 r;  // fma.rn.f32.f16 (FHFMA): see bf16x2_residual
  const unsigned short h0 = (unsigned short)(hi & 0xffffu), h1 = (unsigned short)(hi >> 16);
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r.x) : "h"(h0), "h"((unsigned short)0xBC00), "f"(x0));  // 0xBC00 = -1.0
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r.y) : "h"(h1), "h"((unsigned short)0xBC00), "f"(x1));
  return r;
}

// "Pinned" variants: volatile asm keeps the program order of MUFU and pack instructions relative to each other, so a
// software-pipelined epilogue (sines of group j+1 issued before the packs of group j) survives the scheduler.
__device__ __forceinline__ float sin_approx_pinned(float x) {
  float r;
  asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2_pinned(float lo, float hi) {
  uint32_t r;
  asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16x2_pinned(float lo, float hi) {
  uint32_t r;
  asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float sin_approx(float x) {
  float r;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float cos_approx(float x) {
  float r;
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// Cody-Waite reduction of x to [-pi, pi] with a two-constant 2*pi, rounding via the 1.5*2^23 trick
// (no F2F/FRND on the XU pipe), so MUFU sees a small argument whatever the size of the pre-activation.
// TWO_CONSTANTS = false drops the correction with 2*pi - fp32(2*pi): an argument error of |n| * 1.75e-7 (|n| <= ~10 turns in
// layer 0), i.e. <= 2e-6 -- used by the precisions whose own error is 1e-5 or more (f16f8, fp16), one FMA less per value.
template <bool TWO_CONSTANTS = true>
__device__ __forceinline__ float reduce_2pi(float x) {
  const float kInv2Pi = 0.15915494309189535f;
  const float kMagic = 12582912.0f;        // 1.5 * 2^23
  const float k2PiHi = 6.2831854820251465f;   // fp32(2*pi)
  const float k2PiLo = -1.7484555e-07f;       // 2*pi - fp32(2*pi)
  float n = __fmaf_rn(x, kInv2Pi, kMagic) - kMagic;
  float r = __fmaf_rn(n, -k2PiHi, x);
  return TWO_CONSTANTS ? __fmaf_rn(n, -k2PiLo, r) : r;
}

}  // namespace ptx
}  // namespace cnf
