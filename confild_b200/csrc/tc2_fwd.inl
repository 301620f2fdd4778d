// Launcher of tc2_forward_kernel for ONE operand precision (CNF_TU_PREC); included by tc2_fwd_<precision>.cu so that
// every precision is its own translation unit (parallel compilation).
#include "tc_plan.cuh"

namespace cnf {
namespace host {
namespace {

template <int PREC, bool STASH, bool PACKED>
int launch_tc2_forward(const FwdArgs& a) {
  static std::atomic<size_t> smem_set[kMaxDevices];
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  const int64_t tiles = tc_num_tiles(a.T, a.P, PACKED ? 1 : 0);
  TcPlan plan;
  if (int rc = make_tc2_plan(di, tiles, &plan)) return rc;
  if (a.query) {
    *a.query = LaunchInfo{plan.grid, kTc2Threads, plan.smem, 1, 512, 2 * kTileM};
    return CNF_OK;
  }
  auto kern = tc2_forward_kernel<PREC, STASH, PACKED>;
  if (int rc = ensure_smem(kern, plan.smem, di.device, smem_set)) return rc;
  kern<<<(unsigned)plan.grid, kTc2Threads, plan.smem, a.stream>>>(a.d, a.packed, a.coords, a.coord_frame_stride, a.shift,
                                                                 a.outs, reinterpret_cast<__half*>(a.stash), a.loss, a.T,
                                                                 a.P, plan.stages);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}

}  // namespace

int CNF_TU_NAME(const FwdArgs& a) {
  if (CNF_TU_PREC == CNF_PREC_F16F8 && a.d.nl > kTc2MaxLayers)
    return fail(CNF_ERR_UNSUPPORTED, "f16f8 supports up to %d hidden layers (got %d)", kTc2MaxLayers, a.d.nl);
  const bool pk = use_packed(a.P) != 0;
  if (a.stash)
    return pk ? launch_tc2_forward<CNF_TU_PREC, true, true>(a) : launch_tc2_forward<CNF_TU_PREC, true, false>(a);
  return pk ? launch_tc2_forward<CNF_TU_PREC, false, true>(a) : launch_tc2_forward<CNF_TU_PREC, false, false>(a);
}

}  // namespace host
}  // namespace cnf
