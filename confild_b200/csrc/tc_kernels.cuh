// Generic tcgen05 / TMEM kernels for the hidden-layer chain of the FiLM-SIREN decoder (sm_100a): H = 256, 384
// (H = 128 has its own TMEM-resident kernels in tc2_kernels.cuh).
//
// One CTA decodes 128-point tiles of one frame at a time; the A operand (activations of the previous layer, bf16 hi/lo
// or fp16) lives in shared memory in the UMMA K-major SWIZZLE_128B layout:
//   warps 0-15  "activation" warps: warp = 4*cg + wq; thread (wq, lane) owns query point wq*32+lane (TMEM lane), column
//               group cg owns columns [cg*H/4, (cg+1)*H/4).  Layer 0 (K = cin) on CUDA cores; per hidden layer
//               tcgen05.ld of the fp32 accumulator 16 columns at a time, + FiLM shift, MUFU sin, split into bf16 hi/lo
//               (or fp16), 16-byte conflict-free st.shared into the next layer's A operand; last layer: partial dot
//               with the output head, the four column groups meet in shared memory.
//   warps 16-18 MMA issuers: each whole warp walks the schedule, one elected lane issues tcgen05.mma (M=128, K=16);
//               accumulators in TMEM.  The forward kernels for H=256/384 use one warp per 128-column accumulator block
//               (see kBlockPipe); the backward kernels and H=128 use warp 16 only.
//   warp 19     one lane streams the pre-swizzled 16 KiB weight stages from L2 with 1-D bulk TMA copies.
// Hand-shakes: b_full/b_empty per ring slot, a_full (A operand written, 512 arrivals), d_full (layer's MMAs done).
// Sixteen activation warps (four per SM sub-partition) bring the epilogue to the MUFU bound; MMA and epilogue of a tile
// do not overlap here (a second 128 x H operand does not fit: 2 x 192 KB at H = 384).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <type_traits>

#include "layout.cuh"
#include "ptx.cuh"
#include "tc_common.cuh"

namespace cnf {

constexpr int kTcEpiWarps = 16;
constexpr int kTcIssuerWarps = 3;
constexpr int kTcThreads = (kTcEpiWarps + kTcIssuerWarps + 1) * 32;
constexpr int kTcTailBytes = 384;  // shared-memory area reserved for TcSmemTail

template <int H, int PREC>
struct TcCfg {
  static constexpr bool kSplit = (PREC == CNF_PREC_BF16X3);
  // f16f8: part 0 = fp16 operands, part 1 = the fp8 operands (A: [e5m2(a - fp16 a) | e5m2(a)], B: [e4m3(c S w) |
  // e4m3(c S (w - fp16 w))] per K slab, 128 bytes per row -- the same bytes as a 16-bit slab, so every size below holds;
  // see f16f8_operands16 in tc_common.cuh)
  static constexpr bool kF8 = (PREC == CNF_PREC_F16F8);
  static constexpr int kParts = (kSplit || kF8) ? 2 : 1;
  static constexpr int kSlabs = H / kSlabK;
  static constexpr int kNBlocks = H / kStageRows;
  // The first kATmemBlocks 128-column blocks of the A operand live in the TMEM columns the accumulator leaves free
  // ([H, 512): packed hi 64 columns + lo 64 columns per block) and are read by the MMA from TMEM (85 clk per MMA
  // instead of 116 from shared memory); the rest of A is in shared memory.  H=256: all of A; H=384: one third.
  static constexpr int kATmemBlocks = ((512 - H) / 128) < (H / 128) ? ((512 - H) / 128) : (H / 128);
  static constexpr int kATmemCols = kATmemBlocks * 128;       // activation columns (= K range) resident in TMEM
  static constexpr int kASmemSlabs = kSlabs - 2 * kATmemBlocks;
  static constexpr int kAPartBytes = kASmemSlabs * kTileM * 128;  // shared-memory part of one 16-bit operand
  static constexpr int kABytes = kParts * kAPartBytes;
  static constexpr int kStagesPerLayer = kNBlocks * kSlabs * kParts;
  static constexpr int kColsPerGroup = H / 4;  // columns per activation column group
  // Forward block pipeline (H = 256, 384; NB = H/128 accumulator blocks = K parts).  A layer's MMAs are issued as NB x NB
  // quadrants Q(n,k) = accumulator block n (N=128) x K part k (two K slabs = activation columns [128k, 128k+128)), in the
  // order k-major: Q(0,0) Q(1,0) .. Q(0,1) Q(1,1) ..  Block n is complete after Q(n,NB-1), i.e. BEFORE the blocks after it,
  // so its epilogue E_n (which writes K part n of the next layer's A operand) runs under the remaining quadrants, and the
  // next layer's Q'(n,0) may start as soon as part 0 is written and block n is drained -- under E_1.., so that MMA and
  // epilogue of a single tile overlap almost completely (no second tile fits: TMEM and shared memory are full).
  static constexpr bool kBlockPipe = (H == 256 || H == 384);
  static constexpr uint32_t kTmemCols = (H + kATmemCols) <= 256 ? 256u : 512u;
  static constexpr uint32_t kIdesc = ptx::make_idesc_f16(kSplit ? 1u : 0u, kTileM, kStageRows);
  static_assert(kColsPerGroup % 16 == 0, "column groups are processed 16 columns at a time");
  static_assert(!kF8 || kBlockPipe, "f16f8 is implemented for the block-pipelined widths (H = 256, 384) here; H = 128 has tc2");
};

struct TcSmemTail {  // lives after the A operand and the weight ring
  uint64_t b_full[kTcMaxStages];
  uint64_t b_empty[kTcMaxStages];
  uint64_t a_full;
  uint64_t d_full;
  // kBlockPipe (forward, H = 256/384):
  uint64_t e_done[3];     // K part n of the next A operand written (and accumulator block n drained); 512 arrivals
  uint64_t d_drained[3];  // accumulator block n is in registers (n >= 1); 512 arrivals
  uint64_t d_done[3];     // accumulator block n complete: commit of issuer warp n after Q(n, NB-1)
  uint64_t turn[3];       // issue token passed round-robin between the issuer warps
  uint32_t tmem_base;
};

// Cluster modes of the block-pipelined forward kernel (template parameter CM):
//   kClusterNone   single CTAs
//   kClusterMcast  2-CTA clusters, different tiles, every weight stage fetched once per pair and MULTICAST into both rings
//   kClusterPair   2-CTA clusters driven by cta_group::2 MMAs issued from the leader CTA: M = 256 = the two CTAs' tiles,
//                  the weight operand split along N, so each CTA stores and reads only HALF of every weight stage --
//                  which is what relieves the H = 384 kernel's bound, shared-memory bandwidth
constexpr int kClusterNone = 0, kClusterMcast = 1, kClusterPair = 2;
__host__ __device__ constexpr int tc_slot_bytes(int cm) { return cm == kClusterPair ? kStageBytes / 2 : kStageBytes; }

template <int H, int PREC>
__host__ __device__ constexpr size_t tc_smem_bytes(int num_stages, int slot_bytes = kStageBytes) {
  static_assert(sizeof(TcSmemTail) <= kTcTailBytes, "TcSmemTail outgrew its reserved area");
  return 1024 /*alignment slack*/ + (size_t)TcCfg<H, PREC>::kABytes + (size_t)num_stages * slot_bytes + kTcTailBytes +
         (TcCfg<H, PREC>::kABytes >= 4 * kTileM * 16 ? 0 : 4 * kTileM * 16) + (size_t)H * 4 /* staged FiLM shifts */;
}

// ------------------------------------------------------------------ shared pieces
// Issue all MMAs of one hidden layer.  Called by the WHOLE (converged) MMA warp: every lane polls the barriers and one
// elected lane issues, which keeps the tcgen05 operands in uniform registers (a divergent single-lane loop pays R2UR
// moves and a waterfall loop per MMA and cannot keep the tensor pipe fed).  `slot`/`phase` walk the weight ring.
template <int H, int PREC>
__device__ __forceinline__ void tc_issue_slabs(int ks_begin, int ks_end, uint32_t a_addr, uint32_t ring_addr,
                                               uint32_t tmem_d, TcSmemTail* tail, int num_stages, int& slot,
                                               uint32_t& phase) {
  using C = TcCfg<H, PREC>;
  constexpr int NB = C::kNBlocks;
  // One weight "group" = the NB adjacent ring slots holding all 128-row blocks of one (K slab, hi/lo part).  Row blocks
  // 0 and 1 are contiguous in shared memory, so they feed ONE N=256 MMA (128 clk instead of 2 x 66-72); a third block
  // (H=384) gets an N=128 MMA into the next accumulator columns, which also interleaves independent accumulators.
  constexpr uint32_t kIdescWide = ptx::make_idesc_f16(C::kSplit ? 1u : 0u, kTileM, NB >= 2 ? 256 : 128);
  constexpr uint32_t kIdescTail = ptx::make_idesc_f16(C::kSplit ? 1u : 0u, kTileM, 128);
#pragma unroll 1
  for (int ks = ks_begin; ks < ks_end; ++ks) {
    const bool a_in_tmem = ks < 2 * C::kATmemBlocks;
    // A from TMEM: K slab ks = columns [64*ks, 64*ks+64) of block ks/2 -> 8 packed columns per K=16 step
    const uint32_t at_hi = tmem_d + H + (ks / 2) * 128 + (ks & 1) * 32;
    const int ss = a_in_tmem ? 0 : ks - 2 * C::kATmemBlocks;
    const uint64_t a_hi = ptx::make_desc_k_sw128(a_addr + ss * (kTileM * 128));
    const uint64_t a_lo = ptx::make_desc_k_sw128(a_addr + C::kAPartBytes + ss * (kTileM * 128));
#pragma unroll
    for (int part = 0; part < C::kParts; ++part) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {  // the group's slots never wrap: num_stages is a multiple of NB
        ptx::mbar_wait(&tail->b_full[slot + j], phase);
      }
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint64_t b = ptx::make_desc_k_sw128(ring_addr + slot * kStageBytes);
        const uint64_t b2 = ptx::make_desc_k_sw128(ring_addr + (slot + 2) * kStageBytes);  // third row block (NB == 3)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {  // 4 x K=16 inside the 128-byte row: +32 bytes = +2 in the address field
          const uint32_t first = (part == 0) ? (uint32_t)((ks | kk) != 0) : 1u;
          if (a_in_tmem) {
            ptx::umma_f16_ts(tmem_d, at_hi + kk * 8, b + 2 * kk, kIdescWide, first);
            if (NB == 3) ptx::umma_f16_ts(tmem_d + 256, at_hi + kk * 8, b2 + 2 * kk, kIdescTail, first);
            if (C::kSplit && part == 0) {
              ptx::umma_f16_ts(tmem_d, at_hi + 64 + kk * 8, b + 2 * kk, kIdescWide, 1u);
              if (NB == 3) ptx::umma_f16_ts(tmem_d + 256, at_hi + 64 + kk * 8, b2 + 2 * kk, kIdescTail, 1u);
            }
          } else {
            ptx::umma_f16_ss(tmem_d, a_hi + 2 * kk, b + 2 * kk, kIdescWide, first);
            if (NB == 3) ptx::umma_f16_ss(tmem_d + 256, a_hi + 2 * kk, b2 + 2 * kk, kIdescTail, first);
            if (C::kSplit && part == 0) {
              ptx::umma_f16_ss(tmem_d, a_lo + 2 * kk, b + 2 * kk, kIdescWide, 1u);
              if (NB == 3) ptx::umma_f16_ss(tmem_d + 256, a_lo + 2 * kk, b2 + 2 * kk, kIdescTail, 1u);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) ptx::umma_commit(&tail->b_empty[slot + j]);
      }
      __syncwarp();
      slot += NB;
      if (slot >= num_stages) { slot = 0; phase ^= 1u; }
    }
  }
}

template <int H, int PREC>
__device__ __forceinline__ void tc_issue_layer(uint32_t a_addr, uint32_t ring_addr, uint32_t tmem_d, TcSmemTail* tail,
                                               int num_stages, int& slot, uint32_t& phase) {
  tc_issue_slabs<H, PREC>(0, TcCfg<H, PREC>::kSlabs, a_addr, ring_addr, tmem_d, tail, num_stages, slot, phase);
  if (ptx::elect_one()) ptx::umma_commit(&tail->d_full);
  __syncwarp();
}

// Convert 16 activations (columns c0..c0+15 of this thread's row) to the 16-bit operand format and store them into
// the A operand (K-major SWIZZLE_128B): 2 chunks of 16 bytes per part.  Eight consecutive rows hit eight different
// 16-byte chunk positions, so the stores are bank-conflict free.
template <int H, int PREC>
__device__ __forceinline__ void tc_store_a16(uint8_t* a_smem, uint32_t tmem_row, int row, int c0,
                                             const float (&h)[16]) {
  using C = TcCfg<H, PREC>;
  if constexpr (C::kF8) {
    uint32_t hi[8], lo8[4], a8[4];
    f16f8_operands16(h, hi, lo8, a8);
    if (c0 < C::kATmemCols) {  // TMEM-resident 128-column block: fp16 in [0,64), fp8 operand of slab s in [64 + 32s, +32)
      const uint32_t tb = tmem_row + H + (c0 / 128) * 128;
      const int cb = c0 % 128, slab = cb >> 6, k0 = cb & 63;
      ptx::tmem_st_32x32b_x8(tb + cb / 2, hi);
      ptx::tmem_st_32x32b_x4(tb + 64 + slab * 32 + k0 / 4, lo8);
      ptx::tmem_st_32x32b_x4(tb + 64 + slab * 32 + 16 + k0 / 4, a8);
      return;
    }
    const int cs = c0 - C::kATmemCols;
    uint8_t* rowp = a_smem + (cs / kSlabK) * (kTileM * 128) + row * 128;
    const uint32_t k0 = cs % kSlabK;
    const uint32_t x = row & 7;
    *reinterpret_cast<uint4*>(rowp + (((k0 / 8) ^ x) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(rowp + (((k0 / 8 + 1) ^ x) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    uint8_t* row8 = rowp + C::kAPartBytes;  // the fp8 operand's row: 16-byte chunk k0/16 (a_lo) and 4 + k0/16 (a)
    *reinterpret_cast<uint4*>(row8 + (((k0 / 16) ^ x) << 4)) = make_uint4(lo8[0], lo8[1], lo8[2], lo8[3]);
    *reinterpret_cast<uint4*>(row8 + (((4 + k0 / 16) ^ x) << 4)) = make_uint4(a8[0], a8[1], a8[2], a8[3]);
    return;
  }
  if (c0 < C::kATmemCols) {  // this K range of the operand is TMEM-resident (warp-uniform branch)
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float x0 = h[2 * e], x1 = h[2 * e + 1];
      if (C::kSplit) {
        hi[e] = ptx::pack_bf16x2(x0, x1);
        const float2 r = ptx::bf16x2_residual(hi[e], x0, x1);
        lo[e] = ptx::pack_bf16x2(r.x, r.y);
      } else {
        hi[e] = ptx::pack_f16x2(x0, x1);
      }
    }
    const uint32_t ta = tmem_row + H + (c0 / 128) * 128 + (c0 % 128) / 2;
    ptx::tmem_st_32x32b_x8(ta, hi);
    if (C::kSplit) ptx::tmem_st_32x32b_x8(ta + 64, lo);
    return;
  }
  const int cs = c0 - C::kATmemCols;
  uint8_t* rowp = a_smem + (cs / kSlabK) * (kTileM * 128) + row * 128;
  const uint32_t cbase = (cs % kSlabK) / 8;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x0 = h[q * 8 + 2 * e], x1 = h[q * 8 + 2 * e + 1];
      if (C::kSplit) {
        hi[e] = ptx::pack_bf16x2(x0, x1);
        const float2 r = ptx::bf16x2_residual(hi[e], x0, x1);
        lo[e] = ptx::pack_bf16x2(r.x, r.y);
      } else {
        hi[e] = ptx::pack_f16x2(x0, x1);
      }
    }
    const uint32_t off = ((cbase + q) ^ (row & 7)) << 4;
    *reinterpret_cast<uint4*>(rowp + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (C::kSplit) *reinterpret_cast<uint4*>(rowp + C::kAPartBytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// Common prologue: barriers, TMEM allocation.  Returns the TMEM base.
template <int H, int PREC, int CM = kClusterNone>
__device__ __forceinline__ uint32_t tc_setup(TcSmemTail* tail, int num_stages, int warp) {
  using C = TcCfg<H, PREC>;
  if (threadIdx.x == 0) {
    const bool leader = CM != kClusterPair || ptx::cluster_ctarank() == 0;
    for (int s = 0; s < num_stages; ++s) {
      // pair: the leader's barrier also counts the relayed arrival of the peer's half of the stage
      ptx::mbar_init(&tail->b_full[s], (CM == kClusterPair && leader) ? 2 : 1);
      ptx::mbar_init(&tail->b_empty[s], CM == kClusterMcast ? 2 : 1);  // multicast: released by the issuers of both CTAs
    }
    ptx::mbar_init(&tail->a_full, kTcEpiWarps * 32);
    ptx::mbar_init(&tail->d_full, 1);
    // pair: one (warp-aggregated) arrival per activation warp of BOTH CTAs, on the leader's barriers
    constexpr uint32_t kEpiArrivals = CM == kClusterPair ? 2 * kTcEpiWarps : kTcEpiWarps * 32;
    for (int n = 0; n < 3; ++n) {
      ptx::mbar_init(&tail->e_done[n], kEpiArrivals);
      ptx::mbar_init(&tail->d_drained[n], kEpiArrivals);
      ptx::mbar_init(&tail->d_done[n], 1);
      ptx::mbar_init(&tail->turn[n], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == kTcEpiWarps) {
    if (CM == kClusterPair) {
      ptx::tmem_alloc_pair(&tail->tmem_base, C::kTmemCols);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(&tail->tmem_base, C::kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CM != kClusterNone) ptx::cluster_sync_all();  // the peer's barriers exist before anything is signalled to them
  ptx::tc_fence_after();
  return tail->tmem_base;
}

// Activation-warp arrival on an issuer-facing barrier (e_done / d_drained).  Single CTA / multicast: every thread
// arrives on its own CTA's barrier.  Pair: the MMAs of both CTAs are issued by the leader, so each warp (after its
// threads' own fences) sends ONE arrival to the LEADER's barrier, release at cluster scope.
template <int CM>
__device__ __forceinline__ void tc_epi_arrive(uint64_t* bar) {
  if (CM != kClusterPair) {
    ptx::mbar_arrive(bar);
  } else {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) ptx::mbar_arrive_cluster(ptx::mapa_shared(ptx::smem_u32(bar), 0));
  }
}

// Activation warps wait for the layer's accumulator: one warp polls the mbarrier, the others sleep on a named barrier.
__device__ __forceinline__ void tc_wait_d_full(TcSmemTail* tail, int warp, uint32_t& d_phase) {
  if (warp == 0) ptx::mbar_wait(&tail->d_full, d_phase);
  d_phase ^= 1u;
  ptx::bar_sync(1, kTcEpiWarps * 32);
  ptx::tc_fence_after();
}

// ------------------------------------------------------------------ forward block pipeline (H = 256, 384)
// Epilogue of accumulator block n for this thread's row: its column group's 32 columns [128n + 32cg, +32) = two
// 16-column groups.  Both groups are loaded at once; for n >= 1 the thread then arrives on d_drained[n] (the issuer may
// overwrite the block with the next layer's Q'(n,0)); sines, bf16 split, store into K part n of the next A operand,
// arrive on e_done[n].  LAST: output head instead of the stores, no arrivals.
template <int H, int PREC, bool LAST, bool STASH, int CM = kClusterNone>
__device__ __forceinline__ void tc_block_epilogue(int n, uint8_t* a_smem, uint32_t tmem_row, int row, int cg,
                                                  const float* __restrict__ shl, const float* __restrict__ w_out,
                                                  int cout, float (&y)[4], __half* stash_l, TcSmemTail* tail,
                                                  float inv) {
  using C = TcCfg<H, PREC>;
  const int c0 = 128 * n + 32 * cg;
  uint32_t v0[16], v1[16];
  float h0[16], h1[16];
  ptx::tmem_ld_32x32b_x16(tmem_row + c0, v0);
  ptx::tmem_ld_32x32b_x16(tmem_row + c0 + 16, v1);
  ptx::tmem_wait_ld();
  if (!LAST && n > 0) {
    ptx::tc_fence_before();
    tc_epi_arrive<CM>(&tail->d_drained[n]);
  }
  tc_sines16<STASH, C::kF8>(v0, shl + c0, h0, STASH ? stash_l + (size_t)c0 * kTileM : nullptr, inv);
  tc_sines16<STASH, C::kF8>(v1, shl + c0 + 16, h1, STASH ? stash_l + (size_t)(c0 + 16) * kTileM : nullptr, inv);
  if (!LAST) {
    tc_store_a16<H, PREC>(a_smem, tmem_row, row, c0, h0);
    tc_store_a16<H, PREC>(a_smem, tmem_row, row, c0 + 16, h1);
    ptx::tmem_wait_st();
    ptx::tc_fence_before();
    if (C::kABytes > 0) ptx::fence_proxy_async_smem();
    tc_epi_arrive<CM>(&tail->e_done[n]);
  } else {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const float(&hh)[16] = g == 0 ? h0 : h1;
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        if (o >= cout) continue;
        float acc = y[o];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(w_out + (size_t)o * H + c0 + g * 16) + q);
          acc = fmaf(w4.x, hh[q * 4 + 0], acc);
          acc = fmaf(w4.y, hh[q * 4 + 1], acc);
          acc = fmaf(w4.z, hh[q * 4 + 2], acc);
          acc = fmaf(w4.w, hh[q * 4 + 3], acc);
        }
        y[o] = acc;
      }
    }
  }
}

// Issuer warp n: the issue episodes (block n, K slab 0..kSlabs-1) of one hidden layer (N=128 MMAs), called as a whole
// converged warp.  The token (turn[]) goes round-robin n -> n+1, which yields the k-major order of kBlockPipe and makes
// consecutive quadrants come from different warps (a warp that has issued MMAs is held until the tensor pipe has taken
// them).  All quadrants of block n come from this warp, so its commit after Q(n, NB-1) covers the whole block; E_n
// additionally relies on the tensor pipe completing MMAs in issue order (the quadrants of other blocks that read K part
// n were issued earlier).  `stage` counts the weight stages consumed so far by ALL warps (ring position).
template <int H, int PREC, int CM = kClusterNone>
__device__ __forceinline__ void tc_issue_block(int n, uint32_t a_addr, uint32_t ring_addr, uint32_t tmem_d,
                                               TcSmemTail* tail, int num_stages, int& slot, uint32_t& phase,
                                               uint32_t& e_phase, uint32_t& turn_phase) {
  using C = TcCfg<H, PREC>;
  constexpr int NB = C::kNBlocks;
  constexpr int kP = C::kParts;  // weight stages of one issue episode: (hi, lo) / (fp16, fp8) / fp16 of one K slab, row block n
  constexpr bool PAIR = (CM == kClusterPair);
  constexpr uint32_t kM = PAIR ? 2 * kTileM : kTileM;  // pair: one instruction covers the tiles of both CTAs
  constexpr uint32_t kIdesc = ptx::make_idesc_f16(C::kSplit ? 1u : 0u, kM, 128);
  constexpr uint32_t kIdescF8 = ptx::make_idesc_f8(ptx::kF8E5M2, ptx::kF8E4M3, kM, 128);
  constexpr int kSlot = tc_slot_bytes(CM);
  // One issue EPISODE = accumulator block n x ONE K slab: all its weight stages are awaited first, then all its MMAs
  // (12 bf16x3 / 8 f16f8 / 4 fp16) leave in a single elected-lane block with the stage releases interleaved, and the
  // token moves on.  Episodes go round-robin over the issuer warps (block 0, 1, .., NB-1 of slab 0, then slab 1, ..), so
  // a warp's barrier waits and descriptor set-up run under the other warps' queued MMAs; one episode per (block, slab,
  // PART) -- the first version -- left the tensor pipe idle ~270 clk between the 4-MMA batches of a warp (100-133 clk
  // per MMA whatever the precision).
  auto skip = [&](int episodes) {
    slot += episodes * kP;
    while (slot >= num_stages) { slot -= num_stages; phase ^= 1u; }
  };
  const uint32_t dcol = tmem_d + n * 128;
  skip(n);  // episodes (0..n-1, slab 0) of the other warps
#pragma unroll 1
  for (int ks = 0; ks < C::kSlabs; ++ks) {
    // operands: K part ks/2 of A written; block n drained before its first episode overwrites it; the slab's stages landed
    if ((ks & 1) == 0) ptx::mbar_wait<PAIR>(&tail->e_done[ks >> 1], e_phase);
    if (ks == 0 && n > 0) ptx::mbar_wait<PAIR>(&tail->d_drained[n], e_phase);
    {
      int sl = slot;
      uint32_t ph = phase;
#pragma unroll
      for (int part = 0; part < kP; ++part) {
        ptx::mbar_wait<PAIR>(&tail->b_full[sl], ph);
        if (++sl >= num_stages) { sl = 0; ph ^= 1u; }
      }
    }
    ptx::mbar_wait(&tail->turn[n], turn_phase);
    turn_phase ^= 1u;
    ptx::tc_fence_after();
    const bool a_in_tmem = ks < 2 * C::kATmemBlocks;
    const uint32_t at_hi = tmem_d + H + (ks / 2) * 128 + (ks & 1) * 32;
    const int ss = a_in_tmem ? 0 : ks - 2 * C::kATmemBlocks;
    const uint64_t a_hi = ptx::make_desc_k_sw128(a_addr + ss * (kTileM * 128));
    const uint64_t a_lo = ptx::make_desc_k_sw128(a_addr + C::kAPartBytes + ss * (kTileM * 128));
    if (ptx::elect_one()) {
      int sl = slot;
#pragma unroll
      for (int part = 0; part < kP; ++part) {
        const uint64_t b = ptx::make_desc_k_sw128(ring_addr + sl * kSlot);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t first = (part == 0) ? (uint32_t)((ks | kk) != 0) : 1u;
          if (C::kF8 && part == 1) {  // fp8 stage: K = 32 per MMA; e5m2(a_lo) x e4m3(S w), then e5m2(a) x e4m3(S w_lo)
            if (a_in_tmem) {
              if (PAIR) ptx::umma_f8_ts_pair(dcol, at_hi + 64 + kk * 8, b + 2 * kk, kIdescF8, 1u);
              else ptx::umma_f8_ts(dcol, at_hi + 64 + kk * 8, b + 2 * kk, kIdescF8, 1u);
            } else {
              if (PAIR) ptx::umma_f8_ss_pair(dcol, a_lo + 2 * kk, b + 2 * kk, kIdescF8, 1u);
              else ptx::umma_f8_ss(dcol, a_lo + 2 * kk, b + 2 * kk, kIdescF8, 1u);
            }
          } else if (a_in_tmem) {
            if (PAIR) {
              ptx::umma_f16_ts_pair(dcol, at_hi + kk * 8, b + 2 * kk, kIdesc, first);
              if (C::kSplit && part == 0) ptx::umma_f16_ts_pair(dcol, at_hi + 64 + kk * 8, b + 2 * kk, kIdesc, 1u);
            } else {
              ptx::umma_f16_ts(dcol, at_hi + kk * 8, b + 2 * kk, kIdesc, first);
              if (C::kSplit && part == 0) ptx::umma_f16_ts(dcol, at_hi + 64 + kk * 8, b + 2 * kk, kIdesc, 1u);
            }
          } else {
            if (PAIR) {
              ptx::umma_f16_ss_pair(dcol, a_hi + 2 * kk, b + 2 * kk, kIdesc, first);
              if (C::kSplit && part == 0) ptx::umma_f16_ss_pair(dcol, a_lo + 2 * kk, b + 2 * kk, kIdesc, 1u);
            } else {
              ptx::umma_f16_ss(dcol, a_hi + 2 * kk, b + 2 * kk, kIdesc, first);
              if (C::kSplit && part == 0) ptx::umma_f16_ss(dcol, a_lo + 2 * kk, b + 2 * kk, kIdesc, 1u);
            }
          }
        }
        if (PAIR) ptx::umma_commit_pair(&tail->b_empty[sl], 0x3);             // both CTAs' halves of the stage
        else if (CM == kClusterMcast) ptx::umma_commit_multicast(&tail->b_empty[sl], 0x3);  // the stage sits in both CTAs
        else ptx::umma_commit(&tail->b_empty[sl]);
        if (++sl >= num_stages) sl = 0;
      }
      if (ks == C::kSlabs - 1) {
        if (PAIR) ptx::umma_commit_pair(&tail->d_done[n], 0x3);  // the block is complete in both CTAs
        else ptx::umma_commit(&tail->d_done[n]);
      }
      ptx::mbar_arrive(&tail->turn[n + 1 == NB ? 0 : n + 1]);
    }
    __syncwarp();
    skip(1);                               // this episode
    if (ks + 1 < C::kSlabs) skip(NB - 1);  // the other warps' episodes of this slab and the start of the next
  }
  skip(NB - 1 - n);  // episodes (n+1.., last slab)
  e_phase ^= 1u;
}

// ------------------------------------------------------------------ forward issue stream (round 2)
// The forward kernels issue a layer's episodes (block n x K slab ks) in the order
//     for K part kp:  for block n:  (n, 2kp), (n, 2kp+1)
// i.e. block-major INSIDE a K part.  Block 0's first episodes of the next layer then need only E_0 (K part 0 written,
// block 0 drained) and run under E_1; block 1's need E_1 to have loaded its accumulator, block 2's E_2 -- in the order
// the epilogues run -- and in the last K part block 0 completes four episodes before block 2, so E_0 of the next layer
// starts under the other blocks' MMAs.  (The first version interleaved the blocks slab by slab: (0,ks) (1,ks) (2,ks);
// its third episode of every layer waited for E_2 to START, which stalled everything queued behind it: the trace showed
// a 9.6 k clk hole per layer at H = 384.)  Episodes are dealt round-robin to the three issuer warps along ONE global
// sequence G = 0, 1, 2, ... over all layers and tiles (warp w issues G = w, w+3, ...), so consecutive episodes always
// come from different warps whatever NB; the commit that completes block n is issued by whichever warp holds the block's
// last episode (the tensor pipe completes MMAs in issue order).
template <int H, int PREC, int CM>
__device__ __forceinline__ void tc_issue_stream(int w, int64_t total_layers, uint32_t a_addr, uint32_t ring_addr,
                                                uint32_t tmem_d, TcSmemTail* tail, int num_stages) {
  using C = TcCfg<H, PREC>;
  constexpr int NB = C::kNBlocks;
  constexpr int kP = C::kParts;
  constexpr int kE = C::kSlabs * NB;  // episodes per layer
  constexpr bool PAIR = (CM == kClusterPair);
  constexpr uint32_t kM = PAIR ? 2 * kTileM : kTileM;
  constexpr uint32_t kIdesc = ptx::make_idesc_f16(C::kSplit ? 1u : 0u, kM, 128);
  constexpr uint32_t kIdescF8 = ptx::make_idesc_f8(ptx::kF8E5M2, ptx::kF8E4M3, kM, 128);
  constexpr int kSlot = tc_slot_bytes(CM);
  const int64_t total = total_layers * kE;
  int slot = (w * kP) % num_stages;
  uint32_t phase = (uint32_t)(((w * kP) / num_stages) & 1);
  uint32_t turn_phase = w == 0 ? 1u : 0u;  // warp 0 starts (a fresh barrier passes a parity-1 wait)
  CNF_TRACE_DECL;
  [[maybe_unused]] const bool tracer = (threadIdx.x & 31) == 0;
#pragma unroll 1
  for (int64_t G = w; G < total; G += kTcIssuerWarps) {
    const int64_t layer = G / kE;  // layers issued before this one, over all tiles: parity of the epilogue barriers
    const int idx = (int)(G - layer * kE);
    const int kp = idx / (2 * NB), r = idx - kp * 2 * NB, n = r >> 1, ks = 2 * kp + (r & 1);
    const uint32_t e_phase = (uint32_t)(layer & 1);
    const uint32_t dcol = tmem_d + n * 128;
    if (tracer) CNF_TRACE_EVENT(20 + w, 100000 + layer * 100 + idx);               // episode: start waiting
    ptx::mbar_wait<PAIR>(&tail->e_done[kp], e_phase);                              // K part kp of A written
    if (ks == 0 && n > 0) ptx::mbar_wait<PAIR>(&tail->d_drained[n], e_phase);      // block n in registers
    if (tracer) CNF_TRACE_EVENT(20 + w, 200000 + layer * 100 + idx);               // A operand / accumulator ready
    {
      int sl = slot;
      uint32_t ph = phase;
#pragma unroll
      for (int part = 0; part < kP; ++part) {
        ptx::mbar_wait<PAIR>(&tail->b_full[sl], ph);
        if (++sl >= num_stages) { sl = 0; ph ^= 1u; }
      }
    }
    if (tracer) CNF_TRACE_EVENT(20 + w, 300000 + layer * 100 + idx);  // weight stages landed
    ptx::mbar_wait(&tail->turn[w], turn_phase);
    turn_phase ^= 1u;
    ptx::tc_fence_after();
    if (tracer) CNF_TRACE_EVENT(20 + w, 400000 + layer * 100 + idx);  // issue token received
    const bool a_in_tmem = ks < 2 * C::kATmemBlocks;
    const uint32_t at_hi = tmem_d + H + (ks / 2) * 128 + (ks & 1) * 32;
    const int ss = a_in_tmem ? 0 : ks - 2 * C::kATmemBlocks;
    const uint64_t a_hi = ptx::make_desc_k_sw128(a_addr + ss * (kTileM * 128));
    const uint64_t a_lo = ptx::make_desc_k_sw128(a_addr + C::kAPartBytes + ss * (kTileM * 128));
    if (ptx::elect_one()) {
      int sl = slot;
#pragma unroll
      for (int part = 0; part < kP; ++part) {
        const uint64_t b = ptx::make_desc_k_sw128(ring_addr + sl * kSlot);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t first = (part == 0) ? (uint32_t)((ks | kk) != 0) : 1u;
          if (C::kF8 && part == 1) {  // fp8 stage: K = 32 per MMA; e5m2(a_lo) x e4m3(S w), then e5m2(a) x e4m3(S w_lo)
            if (a_in_tmem) {
              if (PAIR) ptx::umma_f8_ts_pair(dcol, at_hi + 64 + kk * 8, b + 2 * kk, kIdescF8, 1u);
              else ptx::umma_f8_ts(dcol, at_hi + 64 + kk * 8, b + 2 * kk, kIdescF8, 1u);
            } else {
              if (PAIR) ptx::umma_f8_ss_pair(dcol, a_lo + 2 * kk, b + 2 * kk, kIdescF8, 1u);
              else ptx::umma_f8_ss(dcol, a_lo + 2 * kk, b + 2 * kk, kIdescF8, 1u);
            }
          } else if (a_in_tmem) {
            if (PAIR) {
              ptx::umma_f16_ts_pair(dcol, at_hi + kk * 8, b + 2 * kk, kIdesc, first);
              if (C::kSplit && part == 0) ptx::umma_f16_ts_pair(dcol, at_hi + 64 + kk * 8, b + 2 * kk, kIdesc, 1u);
            } else {
              ptx::umma_f16_ts(dcol, at_hi + kk * 8, b + 2 * kk, kIdesc, first);
              if (C::kSplit && part == 0) ptx::umma_f16_ts(dcol, at_hi + 64 + kk * 8, b + 2 * kk, kIdesc, 1u);
            }
          } else {
            if (PAIR) {
              ptx::umma_f16_ss_pair(dcol, a_hi + 2 * kk, b + 2 * kk, kIdesc, first);
              if (C::kSplit && part == 0) ptx::umma_f16_ss_pair(dcol, a_lo + 2 * kk, b + 2 * kk, kIdesc, 1u);
            } else {
              ptx::umma_f16_ss(dcol, a_hi + 2 * kk, b + 2 * kk, kIdesc, first);
              if (C::kSplit && part == 0) ptx::umma_f16_ss(dcol, a_lo + 2 * kk, b + 2 * kk, kIdesc, 1u);
            }
          }
        }
        if (PAIR) ptx::umma_commit_pair(&tail->b_empty[sl], 0x3);
        else if (CM == kClusterMcast) ptx::umma_commit_multicast(&tail->b_empty[sl], 0x3);
        else ptx::umma_commit(&tail->b_empty[sl]);
        if (++sl >= num_stages) sl = 0;
      }
      if (ks == C::kSlabs - 1) {  // the block's last episode: its accumulator is complete
        if (PAIR) ptx::umma_commit_pair(&tail->d_done[n], 0x3);
        else ptx::umma_commit(&tail->d_done[n]);
      }
      ptx::mbar_arrive(&tail->turn[(w + 1) % kTcIssuerWarps]);
    }
    __syncwarp();
    if (tracer) CNF_TRACE_EVENT(20 + w, 500000 + layer * 100 + idx);  // episode issued
    slot += kTcIssuerWarps * kP;  // this warp's next episode is three episodes further along the ring
    while (slot >= num_stages) { slot -= num_stages; phase ^= 1u; }
  }
}

// ------------------------------------------------------------------ forward
// STAGE (block pipeline, frame-aligned tiles only): the layer's FiLM shifts are staged in shared memory once per layer
// instead of being read by every thread with warp-uniform global loads (the H=128 kernel lost 13-20 % without staging).
// CLUSTER (block pipeline only): launched as clusters of two CTAs that decode DIFFERENT tiles but stream the SAME weight
// stages: each CTA fetches half of every 16 KiB stage and multicasts it into both CTAs' rings, so a weight byte is read
// from L2 once per CTA PAIR (the H = 384 kernel is bound by that stream: ~30 B/clk/SM whatever the precision).  A ring
// slot is refilled only after the issuers of BOTH CTAs have released it (b_empty counts two multicast commits).
template <int H, int PREC, bool STASH, bool STAGE = false, int CM = kClusterNone>
__global__ void __launch_bounds__(kTcThreads, 1) tc_forward_kernel(cnf_dims d, const uint8_t* __restrict__ packed,
                                                                   const float* __restrict__ coords,
                                                                   int64_t coord_frame_stride,
                                                                   const float* __restrict__ shift,
                                                                   OutTargets outs, __half* __restrict__ stash,
                                                                   LossArgs loss, int64_t T, int64_t P, int num_stages,
                                                                   int pack_rows) {
  using C = TcCfg<H, PREC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_smem = smem;
  uint8_t* ring = smem + C::kABytes;
  constexpr bool CLUSTER = (CM != kClusterNone);
  constexpr bool PAIR = (CM == kClusterPair);
  constexpr int kSlot = tc_slot_bytes(CM);  // bytes of one ring slot: a whole stage, or this CTA's half of it (pair)
  TcSmemTail* tail = reinterpret_cast<TcSmemTail*>(ring + (size_t)num_stages * kSlot);

  const PackedLayout lay = make_layout(d);
  const int nl = d.nl, cin = d.cin, cout = d.cout;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t PB = (P + kTileM - 1) / kTileM;
  const int64_t tiles = tc_num_tiles(T, P, pack_rows);
  const int64_t SH = (int64_t)(nl + 1) * H;
  static_assert(!CLUSTER || C::kBlockPipe, "the CTA-pair weight multicast is implemented for the block pipeline");
  const uint32_t tmem_base = tc_setup<H, PREC, CM>(tail, num_stages, warp);
  // Tile walk.  Cluster: the pair (even, odd CTA) takes tiles (base, base + 1); both CTAs run the same number of
  // iterations (they share every weight stage), an odd tile count leaves the odd CTA re-decoding the last tile with
  // its results discarded (live == false).
  const uint32_t crank = CLUSTER ? ptx::cluster_ctarank() : 0u;
  const int64_t tile_first = CLUSTER ? (int64_t)(blockIdx.x - crank) : (int64_t)blockIdx.x;
  const int64_t tile_end = CLUSTER ? ((tiles + 1) & ~(int64_t)1) : tiles;
  auto tile_of = [&](int64_t base) { return CLUSTER ? (base + crank < tiles ? base + crank : tiles - 1) : base; };

  if (warp < kTcEpiWarps) {
    // ===================== activation warps =====================
    const int cg = warp / 4, wq = warp % 4;
    const int row = wq * 32 + lane;
    const int col_lo = cg * C::kColsPerGroup, col_hi = col_lo + C::kColsPerGroup;
    const float* w_first_t = reinterpret_cast<const float*>(packed + lay.w_first_t);  // [4][H], coordinate-major
    const float* w_out = reinterpret_cast<const float*>(packed + lay.w_out);
    const float* b_out = reinterpret_cast<const float*>(packed + lay.b_out);
    const uint32_t tmem_row = tmem_base + ((uint32_t)(wq * 32) << 16);
    // [3][128] partial head sums: in the A operand's shared memory when there is one (free after the last layer),
    // else (all of A in TMEM) in a dedicated 6 KiB area behind the barriers
    float4* y_part = C::kABytes >= 4 * kTileM * 16 ? reinterpret_cast<float4*>(a_smem)
                                                   : reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(tail) + kTcTailBytes);
    // one layer of FiLM shifts, behind the barriers and the dedicated head area
    [[maybe_unused]] float* shift_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tail) + kTcTailBytes +
                                                               (C::kABytes >= 4 * kTileM * 16 ? 0 : 4 * kTileM * 16));
    uint32_t d_phase = 0;
    float loss_acc = 0.f;  // fused loss: this thread's share of sum r^2 (head warps: cg == 0)
    CNF_TRACE_DECL;
    const bool tracer = (lane == 0);
    [[maybe_unused]] const int trole = 4 + warp;
    for (int64_t tbase = tile_first; tbase < tile_end; tbase += gridDim.x) {
      const int64_t tile = tile_of(tbase);
      const bool live = !CLUSTER || tbase + crank < tiles;
      const RowMap rm = tc_row_map(tile, row, T, P, PB, pack_rows);
      const int64_t t = rm.t, p = rm.p;
      const bool valid = rm.valid && live;
      const float* sh = shift + t * SH;
      float x[4] = {0.f, 0.f, 0.f, 0.f};
      if (rm.valid) {
        const float* cp = coords + t * coord_frame_stride + p * cin;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < cin) x[j] = cp[j];
      }
      __half* st_row = STASH ? stash + (size_t)tile * SH * kTileM + row * 8 : nullptr;  // tile-major stash

      // ---- layer 0: K = cin on CUDA cores, always range-reduced (|arg| reaches tens of radians)
      if (C::kBlockPipe) {  // no accumulator to drain before the tile's first MMAs (phases must still advance)
        ptx::tc_fence_before();
        for (int n = 1; n < C::kNBlocks; ++n) tc_epi_arrive<CM>(&tail->d_drained[n]);
      }
      // Instantiated per cin, 16-byte loads of the shifts and of the coordinate-major first-layer weights (w_first_t):
      // with a run-time cin and scalar loads this loop was ~19 instructions per element, and it sits in the tile-boundary
      // bubble of the tensor pipe.  Same FMA order as before (shift, coordinate 0, 1, ..): bit-identical results.
      auto layer0 = [&](auto cin_tag) {
      constexpr int CIN = decltype(cin_tag)::value;
#pragma unroll 1
      for (int c = 0; c < C::kColsPerGroup / 16; ++c) {
        // block pipeline: this thread's columns are [128n + 32cg, +32) of every K part n, part 0 first
        const int c0 = C::kBlockPipe ? 128 * (c / 2) + 32 * cg + 16 * (c & 1) : col_lo + c * 16;
        float h[16];
        [[maybe_unused]] float cs[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(sh + c0 + j));
          float z[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
          for (int i = 0; i < CIN; ++i) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(w_first_t + i * H + c0 + j));
            z[0] = fmaf(w4.x, x[i], z[0]);
            z[1] = fmaf(w4.y, x[i], z[1]);
            z[2] = fmaf(w4.z, x[i], z[2]);
            z[3] = fmaf(w4.w, x[i], z[3]);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float r = ptx::reduce_2pi<PREC == CNF_PREC_BF16X3>(z[k]);
            h[j + k] = ptx::sin_approx(r);
            if (STASH) cs[j + k] = ptx::cos_approx(r);
          }
        }
        tc_store_a16<H, PREC>(a_smem, tmem_row, row, c0, h);
        if (STASH) tc_stash16(st_row + (size_t)c0 * kTileM, cs);
        if (C::kBlockPipe && (c & 1)) {  // K part c/2 of the A operand is written
          ptx::tmem_wait_st();
          ptx::tc_fence_before();
          ptx::fence_proxy_async_smem();
          tc_epi_arrive<CM>(&tail->e_done[c / 2]);
        }
      }
      };
      switch (cin) {
        case 1: layer0(std::integral_constant<int, 1>{}); break;
        case 2: layer0(std::integral_constant<int, 2>{}); break;
        case 3: layer0(std::integral_constant<int, 3>{}); break;
        default: layer0(std::integral_constant<int, 4>{}); break;
      }
      if (!C::kBlockPipe) {
        ptx::tmem_wait_st();
        ptx::tc_fence_before();
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&tail->a_full);
      }

      float y[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int l = 1; l <= nl; ++l) {
        const float* shl = sh + (size_t)l * H;
        const bool last = (l == nl);
        if constexpr (C::kBlockPipe) {
          __half* stl = STASH ? st_row + (size_t)l * H * kTileM : nullptr;
          const float inv = C::kF8 ? __ldg(reinterpret_cast<const float*>(packed + lay.tc_scale) + (l - 1)) : 1.f;
          if constexpr (STAGE) {
            ptx::bar_sync(1, kTcEpiWarps * 32);  // nobody still reads the previous layer's shifts
            for (int i = threadIdx.x; i < H; i += kTcEpiWarps * 32) shift_s[i] = __ldg(shl + i);
            // the barrier after the wait on d_done[0] orders these stores before the first read
          }
#pragma unroll 1
          for (int n = 0; n < C::kNBlocks; ++n) {
            // accumulator block n is complete before the blocks after it: its epilogue runs under their MMAs
            if (warp == 0) ptx::mbar_wait(&tail->d_done[n], d_phase);
            ptx::bar_sync(1, kTcEpiWarps * 32);
            ptx::tc_fence_after();
            if (tracer) CNF_TRACE_EVENT(trole, 300 + 10 * l + n);  // block n observed complete
            // two call sites per variant so that each sees a pointer of known address space (ld.shared vs ld.global)
            if constexpr (STAGE) {
              if (!last)
                tc_block_epilogue<H, PREC, false, STASH, CM>(n, a_smem, tmem_row, row, cg, shift_s, w_out, cout, y, stl, tail, inv);
              else
                tc_block_epilogue<H, PREC, true, STASH, CM>(n, a_smem, tmem_row, row, cg, shift_s, w_out, cout, y, stl, tail, inv);
            } else {
              if (!last)
                tc_block_epilogue<H, PREC, false, STASH, CM>(n, a_smem, tmem_row, row, cg, shl, w_out, cout, y, stl, tail, inv);
              else
                tc_block_epilogue<H, PREC, true, STASH, CM>(n, a_smem, tmem_row, row, cg, shl, w_out, cout, y, stl, tail, inv);
            }
            if (tracer) CNF_TRACE_EVENT(trole, 600 + 10 * l + n);  // epilogue of block n done
          }
          d_phase ^= 1u;
          if (last) ptx::tc_fence_before();
        } else {
          tc_wait_d_full(tail, warp, d_phase);
          if (tracer) CNF_TRACE_EVENT(trole, 300 + l);  // d_full observed
#pragma unroll 1
          for (int c0 = col_lo; c0 < col_hi; c0 += 16) {
            uint32_t v[16];
            ptx::tmem_ld_32x32b_x16(tmem_row + c0, v);
            ptx::tmem_wait_ld();
            float h[16];
            tc_sines16<STASH>(v, shl + c0, h, STASH ? st_row + ((size_t)l * H + c0) * kTileM : nullptr);
            if (!last) {
              tc_store_a16<H, PREC>(a_smem, tmem_row, row, c0, h);
            } else {
#pragma unroll
              for (int o = 0; o < 4; ++o) {
                if (o >= cout) continue;
                float acc = y[o];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 w4 = __ldg(reinterpret_cast<const float4*>(w_out + (size_t)o * H + c0) + q);
                  acc = fmaf(w4.x, h[q * 4 + 0], acc);
                  acc = fmaf(w4.y, h[q * 4 + 1], acc);
                  acc = fmaf(w4.z, h[q * 4 + 2], acc);
                  acc = fmaf(w4.w, h[q * 4 + 3], acc);
                }
                y[o] = acc;
              }
            }
          }
          if (!last) {
            ptx::tmem_wait_st();
            ptx::tc_fence_before();
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(&tail->a_full);
          } else {
            ptx::tc_fence_before();
          }
        }
        if (tracer) CNF_TRACE_EVENT(trole, 400 + l);  // epilogue of layer l done
      }
      // ---- head: the four column groups meet in shared memory (the A operand is free until the next tile's layer 0)
      if (cg > 0) y_part[(cg - 1) * kTileM + row] = make_float4(y[0], y[1], y[2], y[3]);
      ptx::bar_sync(1, kTcEpiWarps * 32);
      float* y_stage = reinterpret_cast<float*>(y_part + 3 * kTileM);  // [128][cout] right behind the partial sums
      if (cg == 0) {
        float ys[4] = {y[0], y[1], y[2], y[3]};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float4 yp = y_part[k * kTileM + row];
          ys[0] += yp.x; ys[1] += yp.y; ys[2] += yp.z; ys[3] += yp.w;
        }
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (o < cout) ys[o] += __ldg(b_out + o);
        if (loss.y_meas != nullptr) loss_acc += tc_loss_row(loss, t, p, P, cout, valid, ys);
        if (outs.n == 1) {  // local target: straight from registers
          if (valid && outs.ptr[0] != nullptr) {
            float* op = outs.ptr[0] + (t * P + p) * cout;
#pragma unroll
            for (int o = 0; o < 4; ++o)
              if (o < cout) op[o] = ys[o];
          }
        } else {
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < cout) y_stage[row * cout + o] = ys[o];
        }
      }
      if (outs.n > 1) {  // fused all-gather: contiguous, vectorised stores to every rank's buffer (full sectors on NVLink)
        ptx::bar_sync(1, kTcEpiWarps * 32);
        int64_t q0;
        int nvalid;
        tc_tile_range(tile, T, P, PB, pack_rows, q0, nvalid);
        tc_store_tile(outs, y_stage, q0, live ? nvalid : 0, cout, threadIdx.x, kTcEpiWarps * 32);
      }
      __syncwarp();
      ptx::bar_sync(1, kTcEpiWarps * 32);  // partial sums consumed before the next tile's layer 0 overwrites them
    }
    if (loss.y_meas != nullptr && cg == 0) {  // one slot per head warp: no atomics, deterministic
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, off);
      if (lane == 0) loss.partials[(blockIdx.x * 8 + wq) % kLossPartials] = loss_acc;
    }
    ptx::tc_fence_before();
  } else if (warp < kTcEpiWarps + kTcIssuerWarps) {
    // ===================== MMA issuers (whole warps, one elected lane issues) =====================
    const int which = warp - kTcEpiWarps;
    const uint32_t a_addr = ptx::smem_u32(a_smem), ring_addr = ptx::smem_u32(ring);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    int slot = 0;
    uint32_t b_phase = 0, a_phase = 0;
    CNF_TRACE_DECL;
    if constexpr (C::kBlockPipe) {
      if (PAIR && crank != 0) {
        // peer CTA of a pair: the leader issues the MMAs of both tiles.  One warp here relays "my half of stage s has
        // landed" to the leader's b_full[s] (a bulk copy can only signal a barrier of the CTA it writes to).
        // Three relay warps, stage q handled by warp q % 3 (one warp alone relays a stage every few hundred clocks).
        int64_t my_tiles = 0;
        for (int64_t tbase = tile_first; tbase < tile_end; tbase += gridDim.x) ++my_tiles;
        const int64_t total = my_tiles * nl * C::kStagesPerLayer;
        int rs = which % num_stages;
        uint32_t rph = (uint32_t)((which / num_stages) & 1);
        for (int64_t q = which; q < total; q += kTcIssuerWarps) {
          ptx::mbar_wait(&tail->b_full[rs], rph);
          if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_shared(ptx::smem_u32(&tail->b_full[rs]), 0));
          __syncwarp();
          rs += kTcIssuerWarps;
          while (rs >= num_stages) { rs -= num_stages; rph ^= 1u; }
        }
      } else {  // all three issuer warps share one global episode sequence (tc_issue_stream)
        int64_t my_tiles = 0;
        for (int64_t tbase = tile_first; tbase < tile_end; tbase += gridDim.x) ++my_tiles;
        tc_issue_stream<H, PREC, CM>(which, my_tiles * nl, a_addr, ring_addr, tmem_u, tail, num_stages);
      }
    } else if (which == 0) {
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int l = 1; l <= nl; ++l) {
          ptx::mbar_wait(&tail->a_full, a_phase);
          a_phase ^= 1u;
          if (lane == 0) CNF_TRACE_EVENT(2, 2000 + l);  // A operand ready
          ptx::tc_fence_after();
          tc_issue_layer<H, PREC>(a_addr, ring_addr, tmem_u, tail, num_stages, slot, b_phase);
          if (lane == 0) CNF_TRACE_EVENT(2, 3000 + l);  // the layer's MMAs issued and committed
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== weight producer =====================
    if (lane == 0) {
      const uint8_t* wsrc = packed + (C::kSplit ? lay.tc_fwd_x3 : C::kF8 ? lay.tc_fwd_f8 : lay.tc_fwd_h);
      int slot = 0;
      uint32_t phase = 0;
      for (int64_t tbase = tile_first; tbase < tile_end; tbase += gridDim.x) {
        for (int l = 0; l < nl; ++l) {
          const uint8_t* src = wsrc + (size_t)l * C::kStagesPerLayer * kStageBytes;
          for (int s = 0; s < C::kStagesPerLayer; ++s) {
            // image order: K slab -> part -> row block
            int img = s;
            if (C::kBlockPipe) {  // consumed episode by episode: K part -> row block n -> slab of the part -> part
              constexpr int NB = C::kNBlocks;  // (tc_issue_stream)
              const int e = s / C::kParts, part = s % C::kParts;
              const int kp = e / (2 * NB), r = e % (2 * NB), n = r >> 1, ks = 2 * kp + (r & 1);
              img = (ks * C::kParts + part) * NB + n;
            }
            ptx::mbar_wait(&tail->b_empty[slot], phase ^ 1u);
            constexpr uint32_t kHalf = kStageBytes / 2;
            if (PAIR) {  // this CTA's N-half of the stage (64 weight rows) into its own ring; the MMA reads both CTAs' halves
              ptx::mbar_arrive_expect_tx(&tail->b_full[slot], kHalf);
              ptx::bulk_g2s(ring + (size_t)slot * kHalf, src + (size_t)img * kStageBytes + crank * kHalf, kHalf,
                            &tail->b_full[slot]);
            } else if (CLUSTER) {  // this CTA's half of the stage, delivered to both CTAs (the peer sends the other half)
              ptx::mbar_arrive_expect_tx(&tail->b_full[slot], kStageBytes);
              ptx::bulk_g2s_multicast(ring + (size_t)slot * kStageBytes + crank * kHalf,
                                      src + (size_t)img * kStageBytes + crank * kHalf, kHalf, &tail->b_full[slot], 0x3);
            } else {
              ptx::mbar_arrive_expect_tx(&tail->b_full[slot], kStageBytes);
              ptx::bulk_g2s(ring + (size_t)slot * kStageBytes, src + (size_t)img * kStageBytes, kStageBytes,
                            &tail->b_full[slot]);
            }
            if (++slot == num_stages) { slot = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (CLUSTER) ptx::cluster_sync_all();  // no CTA of the pair exits while the other may still signal its barriers
  if (warp == kTcEpiWarps) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc_pair(tmem_base, C::kTmemCols);
    else ptx::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------ backward (to the FiLM shifts)
// delta_nl = (gout * W_out) .* cos_nl ; for l = nl..1: delta_{l-1} = (delta_l * w0*W_l) .* cos_{l-1};
// gshift[t, l, n] += sum over the tile's points of delta_l[., n].  Always bf16 hi/lo split operands.
template <int H>
__global__ void __launch_bounds__(kTcThreads, 1) tc_backward_kernel(cnf_dims d, const uint8_t* __restrict__ packed,
                                                                    const float* __restrict__ gout,
                                                                    const __half* __restrict__ stash,
                                                                    float* __restrict__ gshift, int64_t T, int64_t P,
                                                                    int num_stages, int pack_rows) {
  constexpr int PREC = CNF_PREC_BF16X3;
  using C = TcCfg<H, PREC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_smem = smem;
  uint8_t* ring = smem + C::kABytes;
  TcSmemTail* tail = reinterpret_cast<TcSmemTail*>(ring + (size_t)num_stages * kStageBytes);

  const PackedLayout lay = make_layout(d);
  const int nl = d.nl, cout = d.cout;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t PB = (P + kTileM - 1) / kTileM;
  const int64_t tiles = tc_num_tiles(T, P, pack_rows);
  const int64_t SH = (int64_t)(nl + 1) * H;
  const uint32_t tmem_base = tc_setup<H, PREC>(tail, num_stages, warp);

  if (warp < kTcEpiWarps) {
    const int cg = warp / 4, wq = warp % 4;
    const int row = wq * 32 + lane;
    const int col_lo = cg * C::kColsPerGroup, col_hi = col_lo + C::kColsPerGroup;
    const float* w_out = reinterpret_cast<const float*>(packed + lay.w_out);
    const uint32_t tmem_row = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint32_t d_phase = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
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
      // ---- seed: delta at the last sine layer
      if (C::kBlockPipe) {  // no accumulator to drain before the tile's first MMAs (phases must still advance)
        ptx::tc_fence_before();
        for (int n = 1; n < C::kNBlocks; ++n) ptx::mbar_arrive(&tail->d_drained[n]);
      }
#pragma unroll 1
      for (int c = 0; c < C::kColsPerGroup / 16; ++c) {
        // block pipeline: this thread's columns are [128n + 32cg, +32) of every K part n, part 0 first
        const int c0 = C::kBlockPipe ? 128 * (c / 2) + 32 * cg + 16 * (c & 1) : col_lo + c * 16;
        float cs[16], dl[16];
        tc_load_cos16(st_row + ((size_t)nl * H + c0) * kTileM, cs);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float g = 0.f;
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < cout) g = fmaf(gy[o], __ldg(w_out + (size_t)o * H + c0 + j), g);
          dl[j] = g * cs[j];
        }
        tc_store_a16<H, PREC>(a_smem, tmem_row, row, c0, dl);
        tc_colsum16_rows(dl, lane, t, gshift + (size_t)nl * H + c0, SH);
        if (C::kBlockPipe && (c & 1)) {  // K part c/2 of the A operand is written
          ptx::tmem_wait_st();
          ptx::tc_fence_before();
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(&tail->e_done[c / 2]);
        }
      }
      if (!C::kBlockPipe) {
        ptx::tmem_wait_st();
        ptx::tc_fence_before();
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&tail->a_full);
      }

#pragma unroll 1
      for (int l = nl; l >= 1; --l) {
        if constexpr (C::kBlockPipe) {
          // block pipeline (see kBlockPipe): accumulator block n = delta columns [128n, 128n+128) is complete before the
          // blocks after it; its epilogue (cos multiply, K part n of the next A operand, column sums) runs under their MMAs
#pragma unroll 1
          for (int n = 0; n < C::kNBlocks; ++n) {
            const int c0 = 128 * n + 32 * cg;
            uint4 cpk[4];  // stashed cos of the layer below for this block's 32 columns: fetched before the wait
#pragma unroll
            for (int q = 0; q < 4; ++q)
              tc_load_cos_chunk(st_row + ((size_t)(l - 1) * H + c0 + q * 8) * kTileM, cpk[q]);
            if (warp == 0) ptx::mbar_wait(&tail->d_done[n], d_phase);
            ptx::bar_sync(1, kTcEpiWarps * 32);
            ptx::tc_fence_after();
            uint32_t v0[16], v1[16];
            ptx::tmem_ld_32x32b_x16(tmem_row + c0, v0);
            ptx::tmem_ld_32x32b_x16(tmem_row + c0 + 16, v1);
            ptx::tmem_wait_ld();
            if (l > 1 && n > 0) {  // block n is in registers: the next layer's Q'(n,0) may overwrite it
              ptx::tc_fence_before();
              ptx::mbar_arrive(&tail->d_drained[n]);
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const uint32_t(&v)[16] = g == 0 ? v0 : v1;
              float cs[16], dl[16];
              tc_unpack_cos16(cpk[2 * g], cpk[2 * g + 1], cs);
#pragma unroll
              for (int j = 0; j < 16; j += 2) {  // mul.f32x2: two columns per instruction
                const float2 m = __fmul2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])),
                                            make_float2(cs[j], cs[j + 1]));
                dl[j] = m.x;
                dl[j + 1] = m.y;
              }
              if (l > 1) tc_store_a16<H, PREC>(a_smem, tmem_row, row, c0 + 16 * g, dl);
              tc_colsum16_rows(dl, lane, t, gshift + (size_t)(l - 1) * H + c0 + 16 * g, SH);
            }
            if (l > 1) {
              ptx::tmem_wait_st();
              ptx::tc_fence_before();
              ptx::fence_proxy_async_smem();
              ptx::mbar_arrive(&tail->e_done[n]);
            }
          }
          d_phase ^= 1u;
          if (l == 1) ptx::tc_fence_before();
        } else {
        // prefetch this thread's stashed cos of the layer below (independent of the MMA) before waiting for it
        constexpr int kChunks = C::kColsPerGroup / 8;
        uint4 cpk[kChunks];
#pragma unroll
        for (int q = 0; q < kChunks; ++q)
          tc_load_cos_chunk(st_row + ((size_t)(l - 1) * H + col_lo + q * 8) * kTileM, cpk[q]);
        tc_wait_d_full(tail, warp, d_phase);
#pragma unroll
        for (int c = 0; c < kChunks / 2; ++c) {
          const int c0 = col_lo + c * 16;
          uint32_t v[16];
          ptx::tmem_ld_32x32b_x16(tmem_row + c0, v);
          float cs[16], dl[16];
          tc_unpack_cos16(cpk[2 * c], cpk[2 * c + 1], cs);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {  // mul.f32x2: two columns per instruction
            const float2 m = __fmul2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])),
                                        make_float2(cs[j], cs[j + 1]));
            dl[j] = m.x;
            dl[j + 1] = m.y;
          }
          if (l > 1) tc_store_a16<H, PREC>(a_smem, tmem_row, row, c0, dl);
          tc_colsum16_rows(dl, lane, t, gshift + (size_t)(l - 1) * H + c0, SH);
        }
        if (l > 1) {
          ptx::tmem_wait_st();
          ptx::tc_fence_before();
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(&tail->a_full);
        } else {
          ptx::tc_fence_before();
        }
        }
      }
    }
  } else if (warp < kTcEpiWarps + kTcIssuerWarps) {
    const int which = warp - kTcEpiWarps;
    const uint32_t a_addr = ptx::smem_u32(a_smem), ring_addr = ptx::smem_u32(ring);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    int slot = 0;
    uint32_t b_phase = 0, a_phase = 0;
    if constexpr (C::kBlockPipe) {
      if (which < C::kNBlocks) {  // issuer warp n owns accumulator block n (see tc_issue_block)
        uint32_t turn_phase = which == 0 ? 1u : 0u;
        for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x)
          for (int l = nl; l >= 1; --l)
            tc_issue_block<H, PREC>(which, a_addr, ring_addr, tmem_u, tail, num_stages, slot, b_phase, a_phase, turn_phase);
      }
    } else if (which == 0) {
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int l = nl; l >= 1; --l) {
          ptx::mbar_wait(&tail->a_full, a_phase);
          a_phase ^= 1u;
          ptx::tc_fence_after();
          tc_issue_layer<H, PREC>(a_addr, ring_addr, tmem_u, tail, num_stages, slot, b_phase);
        }
      }
    }
    __syncwarp();
  } else {
    if (lane == 0) {
      const uint8_t* wsrc = packed + lay.tc_bwd_x3;
      int slot = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int l = nl - 1; l >= 0; --l) {
          const uint8_t* src = wsrc + (size_t)l * C::kStagesPerLayer * kStageBytes;
          for (int s = 0; s < C::kStagesPerLayer; ++s) {
            int img = s;  // block pipeline: quadrant order, see the forward kernel's producer
            if (C::kBlockPipe) {  // consumed episode by episode: K slab -> row block n -> part (tc_issue_block)
              constexpr int NB = C::kNBlocks;
              const int e = s / C::kParts, part = s % C::kParts;
              const int ks = e / NB, n = e % NB;
              img = (ks * C::kParts + part) * NB + n;
            }
            ptx::mbar_wait(&tail->b_empty[slot], phase ^ 1u);
            ptx::mbar_arrive_expect_tx(&tail->b_full[slot], kStageBytes);
            ptx::bulk_g2s(ring + (size_t)slot * kStageBytes, src + (size_t)img * kStageBytes, kStageBytes,
                          &tail->b_full[slot]);
            if (++slot == num_stages) { slot = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == kTcEpiWarps) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

}  // namespace cnf
