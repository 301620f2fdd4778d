// Launcher of tc2_backward_kernel (H = 128, bf16 hi/lo operands).
#include "tc_plan.cuh"

namespace cnf {
namespace host {
namespace {

template <bool PACKED>
int launch_tc2_backward(const BwdArgs& a) {
  static std::atomic<size_t> smem_set[kMaxDevices];
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  const int64_t tiles = tc_num_tiles(a.T, a.P, PACKED ? 1 : 0);
  TcPlan plan;
  if (int rc = make_tc2_plan(di, tiles, &plan)) return rc;
  auto kern = tc2_backward_kernel<PACKED>;
  if (int rc = ensure_smem(kern, plan.smem, di.device, smem_set)) return rc;
  kern<<<(unsigned)plan.grid, kTc2BwdThreads, plan.smem, a.stream>>>(a.d, a.packed, a.gout,
                                                                    reinterpret_cast<const __half*>(a.stash), a.gshift,
                                                                    a.T, a.P, plan.stages);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}

}  // namespace

int tc2_backward(const BwdArgs& a) {
  return use_packed(a.P) ? launch_tc2_backward<true>(a) : launch_tc2_backward<false>(a);
}

}  // namespace host
}  // namespace cnf
CNF_DEFINE_SET_TRACE(set_trace_tc2_bwd)
