// Launch plans of the tensor-core kernels (ring depth, shared memory, grid), shared by forward and backward launchers.
#pragma once
#include "host.cuh"
#include "tc2_kernels.cuh"
#include "tc_kernels.cuh"

namespace cnf {
namespace host {

struct TcPlan {
  int stages = 0;
  size_t smem = 0;
  int64_t grid = 0;
  unsigned tmem_cols = 0;
};

// H = 128 fast path: one CTA per SM, two tiles in flight, the weight ring takes all shared memory (up to 12 stages).
inline int make_tc2_plan(const DeviceInfo& di, int64_t tiles, TcPlan* plan) {
  const size_t fixed = tc2_smem_bytes(0);
  int stages = (int)(((size_t)di.max_smem_optin - fixed) / kStageBytes);
  if (stages > kTcMaxStages) stages = kTcMaxStages;
  const int forced = knobs().stages;
  if (forced >= 4 && forced <= stages) stages = forced;
  if (stages < 6) return fail(CNF_ERR_UNSUPPORTED, "not enough shared memory for the weight ring");
  plan->stages = stages;
  plan->smem = tc2_smem_bytes(stages);
  plan->tmem_cols = 512;
  const int64_t pairs = (tiles + 1) / 2;
  plan->grid = pairs < di.sms ? pairs : di.sms;
  return CNF_OK;
}

// Generic kernels: one CTA per SM (16 activation warps + issuers + producer fill the register file); the weight ring
// takes whatever shared memory the A operand leaves: 6 stages at H=384 (split precisions), 12 otherwise.
template <int H, int PREC>
inline int make_tc_plan(const DeviceInfo& di, int64_t tiles, TcPlan* plan, int slot_bytes = kStageBytes) {
  using C = TcCfg<H, PREC>;
  const size_t fixed = tc_smem_bytes<H, PREC>(0);
  if ((size_t)di.max_smem_optin <= fixed + 2 * kStageBytes)
    return fail(CNF_ERR_UNSUPPORTED, "H=%d precision=%d does not fit in shared memory", H, PREC);
  int stages = (int)(((size_t)di.max_smem_optin - fixed) / slot_bytes);
  if (stages > kTcMaxStages) stages = kTcMaxStages;
  const int forced = knobs().stages;
  if (forced >= 2 && forced <= stages) stages = forced;
  stages -= stages % C::kNBlocks;  // the MMA warp consumes the ring in groups of kNBlocks adjacent slots
  if (stages < C::kNBlocks) return fail(CNF_ERR_UNSUPPORTED, "weight ring too small for H=%d", H);
  plan->stages = stages;
  plan->smem = tc_smem_bytes<H, PREC>(stages, slot_bytes);
  plan->tmem_cols = C::kTmemCols;
  plan->grid = tiles < di.sms ? tiles : di.sms;
  return CNF_OK;
}

}  // namespace host
}  // namespace cnf
