// Launcher of tc3_forward_kernel (H = 128, three tile slots per SM) for the f16f8 precision.
#include "tc_plan.cuh"
#include "tc3_kernels.cuh"

CNF_DEFINE_SET_TRACE(set_trace_tc3_fwd_f16f8)

namespace cnf {
namespace host {

int tc3_forward_f16f8(const FwdArgs& a) {
  static std::atomic<size_t> smem_set[kMaxDevices];
  if (a.d.nl > kTc2MaxLayers) return fail(CNF_ERR_UNSUPPORTED, "f16f8 supports up to %d hidden layers", kTc2MaxLayers);
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  const int64_t tiles = tc_num_tiles(a.T, a.P, 0);
  int stages = (int)(((size_t)di.max_smem_optin - tc3_smem_bytes(0)) / kStageBytes);
  if (stages > kTcMaxStages) stages = kTcMaxStages;
  if (stages < kTc3MinStages) return fail(CNF_ERR_UNSUPPORTED, "not enough shared memory for the three-slot kernel");
  const size_t smem = tc3_smem_bytes(stages);
  const int64_t trips = (tiles + kTc3Slots - 1) / kTc3Slots;
  const int64_t grid = trips < di.sms ? trips : di.sms;
  if (a.query) {
    *a.query = LaunchInfo{grid, kTc3Threads, smem, 1, 512, kTc3Slots * kTileM};
    return CNF_OK;
  }
  auto kern = tc3_forward_kernel<CNF_PREC_F16F8>;
  if (int rc = ensure_smem(kern, smem, di.device, smem_set)) return rc;
  kern<<<(unsigned)grid, kTc3Threads, smem, a.stream>>>(a.d, a.packed, a.coords, a.coord_frame_stride, a.shift,
                                                       a.outs.ptr[0], a.loss, a.T, a.P, stages);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}

}  // namespace host
}  // namespace cnf
