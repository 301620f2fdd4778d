// Launcher of tc_backward_kernel (H = 128 / 256 / 384, bf16 hi/lo operands).
#include "tc_plan.cuh"

namespace cnf {
namespace host {
namespace {

template <int H>
int launch_tc_backward(const BwdArgs& a) {
  static std::atomic<size_t> smem_set[kMaxDevices];
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  const int pack_rows = use_packed(a.P);
  const int64_t tiles = tc_num_tiles(a.T, a.P, pack_rows);
  TcPlan plan;
  if (int rc = make_tc_plan<H, CNF_PREC_BF16X3>(di, tiles, &plan)) return rc;
  auto kern = tc_backward_kernel<H>;
  if (int rc = ensure_smem(kern, plan.smem, di.device, smem_set)) return rc;
  kern<<<(unsigned)plan.grid, kTcThreads, plan.smem, a.stream>>>(a.d, a.packed, a.gout,
                                                                reinterpret_cast<const __half*>(a.stash), a.gshift, a.T,
                                                                a.P, plan.stages, pack_rows);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}

}  // namespace

int tc_backward(const BwdArgs& a) {
  switch (a.d.H) {
    case 128: return launch_tc_backward<128>(a);
    case 256: return launch_tc_backward<256>(a);
    case 384: return launch_tc_backward<384>(a);
  }
  return fail(CNF_ERR_UNSUPPORTED, "no tensor-core kernel for H=%d", a.d.H);
}

}  // namespace host
}  // namespace cnf
CNF_DEFINE_SET_TRACE(set_trace_tc_bwd)
