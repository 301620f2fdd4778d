// Launcher of tc_forward_kernel (H = 128 / 256 / 384) for ONE operand precision (CNF_TU_PREC); included by
// tc_fwd_<precision>.cu so that every precision is its own translation unit.
#include "tc_plan.cuh"

namespace cnf {
namespace host {
namespace {

template <int H, int PREC, bool STASH>
int launch_tc_forward(const FwdArgs& a) {
  static std::atomic<size_t> smem_set[6][kMaxDevices];
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  const int pack_rows = use_packed(a.P);
  const int64_t tiles = tc_num_tiles(a.T, a.P, pack_rows);
  TcPlan plan;
  if (int rc = make_tc_plan<H, PREC>(di, tiles, &plan)) return rc;
  if (a.query) {
    *a.query = LaunchInfo{plan.grid, kTcThreads, plan.smem, 1, (int)plan.tmem_cols, kTileM};
    return CNF_OK;
  }
  // frame-aligned tiles of the block-pipelined kernels stage the layer's FiLM shifts in shared memory
  constexpr bool kCanStage = TcCfg<H, PREC>::kBlockPipe;
  const bool stage = kCanStage && !pack_rows;
  // CTA pairs (2-CTA clusters) that share the weight stream: worth it once there are at least two tiles.
  //   CNF_TC_CLUSTER = 1 (default): every stage multicast into both rings;  2: cta_group::2 MMAs, each CTA holds half of
  //   every stage (works; ~10 % slower end to end than the multicast pair at H = 384: kept for further work)
  constexpr bool kCanCluster = TcCfg<H, PREC>::kBlockPipe;
  const int cm = (kCanCluster && tiles >= 2) ? knobs().cluster : 0;
  if (cm == kClusterMcast || cm == kClusterPair) {
    if (cm == kClusterPair)
      if (int rc = make_tc_plan<H, PREC>(di, tiles, &plan, tc_slot_bytes(kClusterPair))) return rc;
    auto kern = cm == kClusterPair
                    ? (stage ? tc_forward_kernel<H, PREC, STASH, kCanStage, kCanCluster ? kClusterPair : 0>
                             : tc_forward_kernel<H, PREC, STASH, false, kCanCluster ? kClusterPair : 0>)
                    : (stage ? tc_forward_kernel<H, PREC, STASH, kCanStage, kCanCluster ? kClusterMcast : 0>
                             : tc_forward_kernel<H, PREC, STASH, false, kCanCluster ? kClusterMcast : 0>);
    if (int rc = ensure_smem(kern, plan.smem, di.device, smem_set[2 * cm + (stage ? 1 : 0)])) return rc;
    int64_t grid = (tiles + 1) & ~(int64_t)1;
    const int64_t cap = di.sms & ~1;
    if (grid > cap) grid = cap;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = a.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CNF_CUDA(cudaLaunchKernelEx(&cfg, kern, a.d, a.packed, a.coords, a.coord_frame_stride, a.shift, a.outs,
                                reinterpret_cast<__half*>(a.stash), a.loss, a.T, a.P, plan.stages, pack_rows));
    return CNF_OK;
  }
  auto kern = stage ? tc_forward_kernel<H, PREC, STASH, kCanStage> : tc_forward_kernel<H, PREC, STASH, false>;
  if (int rc = ensure_smem(kern, plan.smem, di.device, smem_set[stage ? 1 : 0])) return rc;
  kern<<<(unsigned)plan.grid, kTcThreads, plan.smem, a.stream>>>(a.d, a.packed, a.coords, a.coord_frame_stride, a.shift,
                                                                a.outs, reinterpret_cast<__half*>(a.stash), a.loss, a.T,
                                                                a.P, plan.stages, pack_rows);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}

template <int H>
int dispatch_stash(const FwdArgs& a) {
  return a.stash ? launch_tc_forward<H, CNF_TU_PREC, true>(a) : launch_tc_forward<H, CNF_TU_PREC, false>(a);
}

}  // namespace

int CNF_TU_NAME(const FwdArgs& a) {
  switch (a.d.H) {
    case 128:
      if constexpr (CNF_TU_PREC == CNF_PREC_F16F8)
        return fail(CNF_ERR_UNSUPPORTED, "f16f8 at H=128 runs on the TMEM-resident kernels only (CNF_TC2=0 excludes it)");
      else
        return dispatch_stash<128>(a);
    case 256: return dispatch_stash<256>(a);
    case 384: return dispatch_stash<384>(a);
  }
  return fail(CNF_ERR_UNSUPPORTED, "no tensor-core kernel for H=%d", a.d.H);
}

}  // namespace host
}  // namespace cnf
