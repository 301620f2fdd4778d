// H = 128 forward with THREE 128-point tiles in flight per SM (inference only: no stash, frame-aligned tiles, one target).
//
// Why: with two tile slots (tc2_forward_kernel) every resource idles ~40 % of the time in the f16f8 precision -- tensor
// pipe 51 %, MUFU 57 %, issue slots 71 % -- because each slot runs the strictly serial chain
//     epilogue (~1.85 k clk)  ->  second half of the layer's MMAs  ->  commit  ->  epilogue ...
// and two slots are not enough to fill the gaps.  A third slot does not fit in tensor memory next to
// its operands (3 x (128 accumulator + 128 operand columns) > 512), so here
//     TMEM   accumulators of the three slots [0,128) [128,256) [256,384)  +  the A operand of slot 0 in [384,512)
//     SMEM   the A operands of slots 1 and 2 (K-major SWIZZLE_128B, 64 KiB each: fp16 | fp8, or bf16 hi | lo),
//            a 5-deep ring of 16 KiB weight stages, each consumed by all three slots before it is released
// i.e. slot 0's MMAs read A from TMEM (TS), slots 1 and 2 from shared memory (SS).
// The 16 activation warps form two TEAMS of 8 (2 column halves x 4 TMEM lane quarters, as one slot's warps in tc2); the
// epilogue events (tile triple, layer l, slot g) are numbered k = 0, 1, 2, ... in the order (l, g) and team k mod 2 handles
// event k, so consecutive events -- which belong to different slots -- overlap on the two teams while each slot's MMAs
// run under the other two slots' epilogues.  A slot changes teams from layer to layer, so nothing about a tile lives in
// registers across events: (frame, first point) are recomputed from the tile index at every event.
#pragma once
#include <cuda_runtime.h>

#include "layout.cuh"
#include "ptx.cuh"
#include "tc2_kernels.cuh"
#include "tc_common.cuh"

namespace cnf {

constexpr int kTc3Slots = 3;
constexpr int kTc3Threads = kTc2Threads;          // 16 activation warps + two MMA issuer warps + weight producer
constexpr int kTc3ASlotBytes = 4 * kStageBytes;   // one shared-memory operand: 2 K slabs x 2 parts x 16 KiB
constexpr int kTc3MinStages = 5;

struct Tc3SmemTail {
  float shift_s[kTc3Slots][kTc2H];     // the layer's FiLM shifts of the slot's frame (written after d_full is observed)
  float y_part[kTc3Slots][kTileM][4];  // head partial sums of the upper column half
  float w_first_s[kTc2H * 4];
  float w_out_s[4 * kTc2H];
  uint64_t b_full[kTcMaxStages];
  uint64_t b_empty[kTcMaxStages];      // 3 arrivals: one per slot
  uint64_t a_half[kTc3Slots];          // K slab 0 of the slot's A operand written and its accumulator drained (256 arrivals)
  uint64_t a_full[kTc3Slots];
  uint64_t d_full[kTc3Slots];          // the layer's MMAs of the slot done (one commit per issuer warp)
  uint64_t d_free[kTc3Slots];          // the slot's tile is finished (head done): its accumulator / operand / buffers are free
  uint64_t turn[2];
  uint32_t tmem_base;
  float inv_scale[kTc2MaxLayers];
};

__host__ __device__ constexpr size_t tc3_smem_bytes(int num_stages) {
  return 1024 + (size_t)num_stages * kStageBytes + 2 * (size_t)kTc3ASlotBytes + sizeof(Tc3SmemTail);
}

// 16 activations (columns c0..c0+15 of row `row`) into a shared-memory A operand (same formats as tc2_store_a16).
template <int PREC>
__device__ __forceinline__ void tc3_store_a16_smem(uint8_t* a_smem, int row, int c0, const float (&h)[16]) {
  constexpr int kPart = 2 * kStageBytes;  // bytes of one part (two K slabs)
  const int slab = c0 >> 6, k0 = c0 & 63;
  uint8_t* rowp = a_smem + slab * kStageBytes + row * 128;
  const uint32_t x = row & 7;
  if constexpr (PREC == CNF_PREC_F16F8) {
    uint32_t hi[8], lo8[4], a8[4];
    f16f8_operands16(h, hi, lo8, a8);
    *reinterpret_cast<uint4*>(rowp + (((k0 / 8) ^ x) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(rowp + (((k0 / 8 + 1) ^ x) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    *reinterpret_cast<uint4*>(rowp + kPart + (((k0 / 16) ^ x) << 4)) = make_uint4(lo8[0], lo8[1], lo8[2], lo8[3]);
    *reinterpret_cast<uint4*>(rowp + kPart + (((4 + k0 / 16) ^ x) << 4)) = make_uint4(a8[0], a8[1], a8[2], a8[3]);
  } else {
    constexpr bool kSplit = (PREC == CNF_PREC_BF16X3);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float x0 = h[q * 8 + 2 * e], x1 = h[q * 8 + 2 * e + 1];
        if (kSplit) {
          hi[e] = ptx::pack_bf16x2_pinned(x0, x1);
          const float2 r = ptx::bf16x2_residual(hi[e], x0, x1);
          lo[e] = ptx::pack_bf16x2_pinned(r.x, r.y);
        } else {
          hi[e] = ptx::pack_f16x2_pinned(x0, x1);
        }
      }
      const uint32_t off = ((k0 / 8 + q) ^ x) << 4;
      *reinterpret_cast<uint4*>(rowp + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      if (kSplit) *reinterpret_cast<uint4*>(rowp + kPart + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

template <int PREC, bool SMEM_A>
__device__ __forceinline__ void tc3_store_a16(uint32_t tmem_a, uint8_t* a_smem, int row, int c0, const float (&h)[16]) {
  if constexpr (SMEM_A) tc3_store_a16_smem<PREC>(a_smem, row, c0, h);
  else tc2_store_a16<PREC, true>(tmem_a, c0, h);
}
// Make this thread's operand writes visible to the tensor core (async proxy for shared memory, tcgen05 for TMEM).
template <bool SMEM_A>
__device__ __forceinline__ void tc3_publish_a() {
  if constexpr (SMEM_A) {
    ptx::fence_proxy_async_smem();
  } else {
    ptx::tmem_wait_st();
  }
  ptx::tc_fence_before();
}

// One hidden layer for this thread's row and its warpgroup's 64 columns (software pipeline over four 16-column groups as
// in tc2_hidden_layer); the operand goes to TMEM (slot 0) or shared memory (slots 1, 2).
template <int PREC, bool LAST, bool SMEM_A>
__device__ __forceinline__ void tc3_hidden_layer(uint32_t lane_base, uint32_t tmem_a, uint8_t* a_smem, int row, int hf,
                                                 const float* __restrict__ sbuf, const float* __restrict__ w_out_s,
                                                 int cout, float (&y)[4], uint64_t* a_half, uint64_t* a_full, float inv) {
  constexpr bool SCALED = (PREC == CNF_PREC_F16F8);
  uint32_t v[2][16];
  float hcur[16], hnext[16];
  ptx::tmem_ld_32x32b_x16(lane_base + tc2_group_col(hf, 0), v[0]);
  ptx::tmem_wait_ld();
  ptx::tmem_ld_32x32b_x16(lane_base + tc2_group_col(hf, 1), v[1]);
  tc_sines16<false, SCALED>(v[0], sbuf + tc2_group_col(hf, 0), hnext, nullptr, inv);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int c0 = tc2_group_col(hf, c);
#pragma unroll
    for (int j = 0; j < 16; ++j) hcur[j] = hnext[j];
    if (c + 1 < 4) {
      ptx::tmem_wait_ld();
      tc_sines16<false, SCALED>(v[(c + 1) & 1], sbuf + tc2_group_col(hf, c + 1), hnext, nullptr, inv);
      if (c + 2 < 4) ptx::tmem_ld_32x32b_x16(lane_base + tc2_group_col(hf, c + 2), v[c & 1]);
    }
    if (!LAST) {
      tc3_store_a16<PREC, SMEM_A>(tmem_a, a_smem, row, c0, hcur);
      if (c == 1) {
        ptx::tmem_wait_ld();  // group 3 (the last of D) is in registers
        tc3_publish_a<SMEM_A>();
        ptx::mbar_arrive(a_half);
      } else if (c == 3) {
        tc3_publish_a<SMEM_A>();
        ptx::mbar_arrive(a_full);
      }
    } else {
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        if (o >= cout) continue;
        float2 acc = make_float2(y[o], 0.f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w4 = *reinterpret_cast<const float4*>(w_out_s + o * kTc2H + c0 + q * 4);
          acc = __ffma2_rn(make_float2(w4.x, w4.y), make_float2(hcur[q * 4 + 0], hcur[q * 4 + 1]), acc);
          acc = __ffma2_rn(make_float2(w4.z, w4.w), make_float2(hcur[q * 4 + 2], hcur[q * 4 + 3]), acc);
        }
        y[o] = acc.x + acc.y;
      }
    }
  }
}

template <int PREC>
__global__ void __launch_bounds__(kTc3Threads, 1) tc3_forward_kernel(cnf_dims d, const uint8_t* __restrict__ packed,
                                                                     const float* __restrict__ coords,
                                                                     int64_t coord_frame_stride,
                                                                     const float* __restrict__ shift, float* __restrict__ out,
                                                                     LossArgs loss, int64_t T, int64_t P, int num_stages) {
  constexpr int H = kTc2H;
  constexpr bool kSplit = (PREC == CNF_PREC_BF16X3);
  constexpr bool kF8 = (PREC == CNF_PREC_F16F8);
  constexpr int kParts = (kSplit || kF8) ? 2 : 1;
  constexpr int kSPL = (H / kSlabK) * kParts;
  constexpr int kSPH = kSPL / 2;
  constexpr uint32_t kIdesc = ptx::make_idesc_f16(kSplit ? 1u : 0u, kTileM, H);
  [[maybe_unused]] constexpr uint32_t kIdescF8 = ptx::make_idesc_f8(ptx::kF8E5M2, ptx::kF8E4M3, kTileM, H);
  constexpr int kMmaWarp = kTc2EpiWarps;
  constexpr int kPartBytes = 2 * kStageBytes;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_smem_base = ring + (size_t)num_stages * kStageBytes;  // operands of slots 1 and 2
  Tc3SmemTail* tail = reinterpret_cast<Tc3SmemTail*>(a_smem_base + 2 * (size_t)kTc3ASlotBytes);

  const PackedLayout lay = make_layout(d);
  const int nl = d.nl, cin = d.cin, cout = d.cout;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t PB = (P + kTileM - 1) / kTileM;
  const int64_t tiles = T * PB;
  const int64_t trips = (tiles + kTc3Slots - 1) / kTc3Slots;
  const int64_t SH = (int64_t)(nl + 1) * H;

  if (threadIdx.x == 0) {
    for (int s = 0; s < num_stages; ++s) {
      ptx::mbar_init(&tail->b_full[s], 1);
      ptx::mbar_init(&tail->b_empty[s], kTc3Slots);
    }
    for (int g = 0; g < kTc3Slots; ++g) {
      ptx::mbar_init(&tail->a_half[g], 256);
      ptx::mbar_init(&tail->a_full[g], 256);
      ptx::mbar_init(&tail->d_full[g], 2);
      ptx::mbar_init(&tail->d_free[g], 256);
    }
    ptx::mbar_init(&tail->turn[0], 1);
    ptx::mbar_init(&tail->turn[1], 1);
    ptx::fence_mbar_init();
  }
  {
    const float* w_first = reinterpret_cast<const float*>(packed + lay.w_first);
    const float* w_out = reinterpret_cast<const float*>(packed + lay.w_out);
    for (int i = threadIdx.x; i < H * cin; i += kTc3Threads) tail->w_first_s[i] = w_first[i];
    for (int i = threadIdx.x; i < cout * H; i += kTc3Threads) tail->w_out_s[i] = w_out[i];
    if (kF8) {
      const float* sc = reinterpret_cast<const float*>(packed + lay.tc_scale);
      for (int i = threadIdx.x; i < nl && i < kTc2MaxLayers; i += kTc3Threads) tail->inv_scale[i] = sc[i];
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
    // ===================== activation teams =====================
    const int team = warp / 8, hf = (warp / 4) & 1, wq = warp % 4;
    const int row = wq * 32 + lane;
    const int stage_col = 32 * hf + (row & 31) + 64 * ((row >> 5) & 1);  // column whose shift this thread stages (wq < 2)
    const uint32_t bar_wg = 1 + team * 2 + hf;  // 128 threads
    const uint32_t bar_team = 5 + team;         // 256 threads
    const float* b_out = reinterpret_cast<const float*>(packed + lay.b_out);
    float loss_acc = 0.f;
    int64_t it = 0;
    CNF_TRACE_DECL;
    [[maybe_unused]] const bool tracer = (lane == 0 && wq == 0);
    [[maybe_unused]] const int trole = 4 + warp;
    for (int64_t trip = blockIdx.x; trip < trips; trip += gridDim.x, ++it) {
#pragma unroll 1
      for (int l = 0; l <= nl; ++l) {
#pragma unroll
        for (int g = 0; g < kTc3Slots; ++g) {
          const int64_t k = (it * (nl + 1) + l) * kTc3Slots + g;
          const int64_t tile = kTc3Slots * trip + g;
          if (tile >= tiles) continue;  // idle slot of the last triple
          if ((int)(k & 1) != team) {
            // The other team's event.  Its barrier phase is still observed here (by the polling warp), in event order:
            // a parity wait is only meaningful for a waiter that has seen every earlier phase, and with idle slots a team
            // can otherwise fall two phases behind a slot's barrier (the wait for layer l would then return at once on
            // the phase of layer l - 2).  The MMAs of successive events complete in order, so this costs no time.
            if (hf == 0 && wq == 0) {
              if (l > 0) ptx::mbar_wait(&tail->d_full[g], (uint32_t)((it * nl + (l - 1)) & 1));
              else if (it > 0) ptx::mbar_wait(&tail->d_free[g], (uint32_t)((it - 1) & 1));
            }
            continue;
          }
          const int64_t t = (tiles <= 0x7fffffffLL) ? (int64_t)((uint32_t)tile / (uint32_t)PB) : tile / PB;
          const int64_t p0 = (tile - t * PB) * kTileM;
          const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16) + g * 128;
          const uint32_t tmem_a = tmem_base + ((uint32_t)(wq * 32) << 16) + 384;   // slot 0 only
          uint8_t* a_smem = a_smem_base + (g > 0 ? (g - 1) : 0) * (size_t)kTc3ASlotBytes;  // slots 1, 2 only
          if (l == 0) {
            // ---- a new tile for slot g: wait until the slot's previous tile is finished, publish the tile, layer 0
            if (tracer) CNF_TRACE_EVENT(trole, 1000 + l * 10 + g);
            if (it > 0) {
              if (hf == 0 && wq == 0) ptx::mbar_wait(&tail->d_free[g], (uint32_t)((it - 1) & 1));
              ptx::bar_sync(bar_team, 256);
              ptx::tc_fence_after();
            }
            if (tracer) CNF_TRACE_EVENT(trole, 2000 + l * 10 + g);
            const int64_t p = p0 + row;
            const bool valid = p < P;
            const float* sh = shift + t * SH;
            float x[4] = {0.f, 0.f, 0.f, 0.f};
            if (valid) {
              const float* cp = coords + t * coord_frame_stride + p * cin;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j < cin) x[j] = cp[j];
            }
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
#pragma unroll 1
              for (int q = 0; q < 2; ++q) {
                const int c0 = 32 * hf + 64 * half + 16 * q;
                float h[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  float z = __ldg(sh + c0 + j);
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    if (i < cin) z = fmaf(tail->w_first_s[(c0 + j) * cin + i], x[i], z);
                  h[j] = ptx::sin_approx(ptx::reduce_2pi(z));
                }
                if (g == 0) tc3_store_a16<PREC, false>(tmem_a, a_smem, row, c0, h);
                else tc3_store_a16<PREC, true>(tmem_a, a_smem, row, c0, h);
              }
              if (g == 0) tc3_publish_a<false>();
              else tc3_publish_a<true>();
              ptx::mbar_arrive(half == 0 ? &tail->a_half[g] : &tail->a_full[g]);
            }
            if (tracer) CNF_TRACE_EVENT(trole, 4000 + l * 10 + g);
            continue;
          }
          // ---- hidden layer l of slot g
          const bool last = (l == nl);
          const float pre = (wq < 2) ? __ldg(shift + t * SH + (size_t)l * H + stage_col) : 0.f;
          if (tracer) CNF_TRACE_EVENT(trole, 1000 + l * 10 + g);
          if (hf == 0 && wq == 0) ptx::mbar_wait(&tail->d_full[g], (uint32_t)((it * nl + (l - 1)) & 1));
          ptx::bar_sync(bar_team, 256);
          ptx::tc_fence_after();
          if (tracer) CNF_TRACE_EVENT(trole, 2000 + l * 10 + g);
          // the slot's previous layer (handled by the other team) is complete: its shift buffer may be overwritten
          if (wq < 2) tail->shift_s[g][stage_col] = pre;
          ptx::bar_sync(bar_wg, 128);
          const float inv = kF8 ? tail->inv_scale[l - 1] : 1.f;
          float y[4] = {0.f, 0.f, 0.f, 0.f};
          if (!last) {
            if (g == 0)
              tc3_hidden_layer<PREC, false, false>(lane_base, tmem_a, a_smem, row, hf, tail->shift_s[g], tail->w_out_s, cout, y,
                                                   &tail->a_half[g], &tail->a_full[g], inv);
            else
              tc3_hidden_layer<PREC, false, true>(lane_base, tmem_a, a_smem, row, hf, tail->shift_s[g], tail->w_out_s, cout, y,
                                                  &tail->a_half[g], &tail->a_full[g], inv);
            if (tracer) CNF_TRACE_EVENT(trole, 4000 + l * 10 + g);
            continue;
          }
          tc3_hidden_layer<PREC, true, false>(lane_base, tmem_a, a_smem, row, hf, tail->shift_s[g], tail->w_out_s, cout, y,
                                              nullptr, nullptr, inv);
          // ---- head: combine the two column halves, 4*cout bytes per point
          ptx::tc_fence_before();
          if (hf == 1) *reinterpret_cast<float4*>(tail->y_part[g][row]) = make_float4(y[0], y[1], y[2], y[3]);
          ptx::bar_sync(bar_team, 256);
          if (hf == 0) {
            const float4 yp = *reinterpret_cast<const float4*>(tail->y_part[g][row]);
            float ys[4] = {y[0] + yp.x, y[1] + yp.y, y[2] + yp.z, y[3] + yp.w};
#pragma unroll
            for (int o = 0; o < 4; ++o)
              if (o < cout) ys[o] += __ldg(b_out + o);
            const int64_t p = p0 + row;
            const bool valid = p < P;
            if (loss.y_meas != nullptr) loss_acc += tc_loss_row(loss, t, valid ? p : P - 1, P, cout, valid, ys);
            if (valid && out != nullptr) {
              float* op = out + (t * P + p) * cout;
#pragma unroll
              for (int o = 0; o < 4; ++o)
                if (o < cout) op[o] = ys[o];
            }
          }
          __syncwarp();
          if (tracer) CNF_TRACE_EVENT(trole, 4000 + l * 10 + g);
          ptx::mbar_arrive(&tail->d_free[g]);  // accumulator read, y_part consumed, slot bookkeeping no longer needed
        }
      }
    }
    if (loss.y_meas != nullptr && hf == 0) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, off);
      if (lane == 0) loss.partials[(blockIdx.x * 8 + team * 4 + wq) % kLossPartials] = loss_acc;
    }
    ptx::tc_fence_before();
  } else if (warp < kMmaWarp + 2) {
    // ===================== MMA issuers (one warp per K slab, as in tc2_forward_kernel; three slots) =====================
    const int half = warp - kMmaWarp;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t ring_addr = ptx::smem_u32(ring);
    const uint32_t a_addr = ptx::smem_u32(a_smem_base);
    uint32_t turn_phase = half == 0 ? 1u : 0u;
    int slot0 = half * kSPH;
    uint32_t ph0 = 0;
    int64_t it = 0;
    CNF_TRACE_DECL;
    for (int64_t trip = blockIdx.x; trip < trips; trip += gridDim.x, ++it) {
      for (int l = 1; l <= nl; ++l) {
        if (lane == 0) CNF_TRACE_EVENT(2 + half, 500 + l);  // start waiting for the layer's weight stages
        {
          int slot = slot0;
          uint32_t ph = ph0;
#pragma unroll
          for (int s = 0; s < kSPH; ++s) {
            ptx::mbar_wait(&tail->b_full[slot], ph);
            if (++slot >= num_stages) { slot = 0; ph ^= 1u; }
          }
        }
        const uint32_t a_par = (uint32_t)((it * nl + (l - 1)) & 1);
#pragma unroll
        for (int g = 0; g < kTc3Slots; ++g) {
          const bool mine = (kTc3Slots * trip + g < tiles);
          const uint32_t tmem_d = tmem_u + g * 128;
          const uint32_t tmem_a = tmem_u + 384;
          const uint32_t a_g = a_addr + (g > 0 ? (g - 1) : 0) * kTc3ASlotBytes + half * kStageBytes;  // this K slab, part 0
          if (lane == 0) CNF_TRACE_EVENT(2 + half, 1000 + l * 10 + g);  // weights there, waiting for the A operand
          if (mine) ptx::mbar_wait(half == 0 ? &tail->a_half[g] : &tail->a_full[g], a_par);
          if (lane == 0) CNF_TRACE_EVENT(2 + half, 2000 + l * 10 + g);  // A operand ready, waiting for the turn
          ptx::mbar_wait(&tail->turn[half], turn_phase);
          turn_phase ^= 1u;
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            int slot = slot0;
#pragma unroll
            for (int s = 0; s < kSPH; ++s) {
              if (mine) {
                const uint64_t b = ptx::make_desc_k_sw128(ring_addr + slot * kStageBytes);
                const int part = s % kParts;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  const uint32_t first = (part == 0) ? (uint32_t)((half | kk) != 0) : 1u;
                  if (g == 0) {  // A operand in tensor memory
                    const uint32_t a_hi = tmem_a + (half * 4 + kk) * 8;
                    if (part == 0) {
                      ptx::umma_f16_ts(tmem_d, a_hi, b + 2 * kk, kIdesc, first);
                      if (kSplit) ptx::umma_f16_ts(tmem_d, a_hi + 64, b + 2 * kk, kIdesc, 1u);
                    } else if (kF8) {
                      ptx::umma_f8_ts(tmem_d, tmem_a + 64 + half * 32 + kk * 8, b + 2 * kk, kIdescF8, 1u);
                    } else {
                      ptx::umma_f16_ts(tmem_d, a_hi, b + 2 * kk, kIdesc, 1u);
                    }
                  } else {  // A operand in shared memory
                    const uint64_t a0 = ptx::make_desc_k_sw128(a_g);
                    const uint64_t a1 = ptx::make_desc_k_sw128(a_g + kPartBytes);
                    if (part == 0) {
                      ptx::umma_f16_ss(tmem_d, a0 + 2 * kk, b + 2 * kk, kIdesc, first);
                      if (kSplit) ptx::umma_f16_ss(tmem_d, a1 + 2 * kk, b + 2 * kk, kIdesc, 1u);
                    } else if (kF8) {
                      ptx::umma_f8_ss(tmem_d, a1 + 2 * kk, b + 2 * kk, kIdescF8, 1u);
                    } else {
                      ptx::umma_f16_ss(tmem_d, a0 + 2 * kk, b + 2 * kk, kIdesc, 1u);
                    }
                  }
                }
                ptx::umma_commit(&tail->b_empty[slot]);
              } else {
                ptx::mbar_arrive(&tail->b_empty[slot]);  // idle slot: still release its share of the stage
              }
              if (++slot >= num_stages) slot = 0;
            }
            if (mine) ptx::umma_commit(&tail->d_full[g]);
            ptx::mbar_arrive(&tail->turn[half ^ 1]);
          }
          __syncwarp();
          if (lane == 0) CNF_TRACE_EVENT(2 + half, 3000 + l * 10 + g);  // issued
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
      for (int64_t trip = blockIdx.x; trip < trips; trip += gridDim.x) {
        for (int l = 0; l < nl; ++l) {
          const uint8_t* src = wsrc + (size_t)l * kSPL * kStageBytes;
          for (int s = 0; s < kSPL; ++s) {
            ptx::mbar_wait(&tail->b_empty[slot], phase ^ 1u);
            ptx::mbar_arrive_expect_tx(&tail->b_full[slot], kStageBytes);
            ptx::bulk_g2s(ring + (size_t)slot * kStageBytes, src + (size_t)s * kStageBytes, kStageBytes, &tail->b_full[slot]);
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
