// Host-side glue shared by the C-ABI translation unit (cnf_cabi.cu) and the per-kernel launch units
// (tc2_fwd_*.cu, tc2_bwd.cu, tc_fwd_*.cu, tc_bwd.cu): error reporting, cached device attributes, debug knobs read
// once, launch plans and the launcher prototypes.  The library is built from several translation units so that nvcc
// compiles the template-heavy kernels in parallel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/confild_cnf.h"
#include "layout.cuh"
#include "tc_common.cuh"

namespace cnf {
namespace host {

int fail(int code, const char* fmt, ...);  // formats the thread-local error text, returns `code`

#define CNF_CUDA(expr)                                                                                          \
  do {                                                                                                          \
    cudaError_t e__ = (expr);                                                                                   \
    if (e__ != cudaSuccess) return ::cnf::host::fail(CNF_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

constexpr int kMaxDevices = 64;

// Attributes of the current device, queried once per device and process.
struct DeviceInfo {
  int device = 0;
  int sms = 0;
  int max_smem_optin = 0;
};
int device_info(DeviceInfo* info);

// Debug knobs: initial values from the environment, read ONCE per process (never on the call path); tests change
// them through cnf_set_debug_knob.
//   CNF_TC2=0        route H = 128 through the generic kernels        CNF_TC_STAGES=n  cap the weight ring depth
//   CNF_TC_PACKED=0|1 force frame-aligned / packed tiles
struct Knobs {
  int tc2 = 1;
  int stages = 0;
  int packed = -1;
  int cluster = 1;  // CNF_TC_CLUSTER: H = 256/384 forward as CTA pairs: 1 = multicast weight stages (default), 2 = cta_group::2 MMAs, 0 = single CTAs
  int gn_cluster = 256;  // CNF_GN_CLUSTER: GroupNorm of activations up to this many KiB per sample as ONE 8-CTA-cluster launch (0 = always two kernels)
};
const Knobs& knobs();
int set_knob(const char* name, int value);  // tests / tuning: override a knob at run time

// Packed tiles (rows = consecutive (frame, point) pairs, a tile may span frames) when frame-aligned 128-point tiles
// would waste >= 1/6 of their rows on padding: P < 128, or a ragged P of a few hundred.
inline int use_packed(int64_t P) {
  const int forced = knobs().packed;
  if (forced == 0 || forced == 1) return forced;
  const int64_t padded = (P + kTileM - 1) / kTileM * kTileM;
  return padded * 5 >= P * 6 ? 1 : 0;
}

// What a launcher would launch (cnf_query_launch) -- filled instead of launching when FwdArgs::query is set.
struct LaunchInfo {
  int64_t grid = 0;
  int threads = 0;
  size_t smem = 0;
  int ctas_per_sm = 1;
  int tmem_cols = 0;
  int tile_points = 0;
};

struct FwdArgs {
  cnf_dims d;
  const uint8_t* packed;
  const float* coords;
  int64_t coord_frame_stride;
  const float* shift;
  OutTargets outs;
  void* stash;  // nullptr = inference
  LossArgs loss;  // loss.y_meas == nullptr = no fused loss
  int64_t T, P;
  cudaStream_t stream;
  LaunchInfo* query;  // non-null: fill and return without launching
};

struct BwdArgs {
  cnf_dims d;
  const uint8_t* packed;
  const float* gout;
  const void* stash;
  float* gshift;
  int64_t T, P;
  cudaStream_t stream;
};

// One function per (kernel family, precision): each lives in its own translation unit.
int tc2_forward_bf16x3(const FwdArgs& a);
int tc2_forward_fp16(const FwdArgs& a);
int tc2_forward_f16f8(const FwdArgs& a);
int tc2_backward(const BwdArgs& a);
int tc_forward_bf16x3(const FwdArgs& a);
int tc_forward_fp16(const FwdArgs& a);
int tc_forward_f16f8(const FwdArgs& a);
int tc_backward(const BwdArgs& a);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel instantiation, device): `slot` is a static array
// owned by the templated launcher, i.e. unique per kernel instantiation.
template <typename Kernel>
inline int ensure_smem(Kernel kern, size_t smem, int device, std::atomic<size_t>* slot) {
  if (device < 0 || device >= kMaxDevices) device = 0;
  if (slot[device].load(std::memory_order_acquire) >= smem) return CNF_OK;
  CNF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  slot[device].store(smem, std::memory_order_release);
  return CNF_OK;
}

#ifdef CNF_TRACE
int set_trace_tc2_fwd_bf16x3(unsigned long long* p);
int set_trace_tc2_fwd_fp16(unsigned long long* p);
int set_trace_tc2_fwd_f16f8(unsigned long long* p);
int set_trace_tc_fwd_f16f8(unsigned long long* p);
int set_trace_tc2_bwd(unsigned long long* p);
int set_trace_tc_fwd_bf16x3(unsigned long long* p);
int set_trace_tc_fwd_fp16(unsigned long long* p);
int set_trace_tc_bwd(unsigned long long* p);
#define CNF_DEFINE_SET_TRACE(name)                                                              \
  int ::cnf::host::name(unsigned long long* p) {                                                \
    return cudaMemcpyToSymbol(::cnf::g_trace, &p, sizeof(p)) == cudaSuccess ? 0 : CNF_ERR_CUDA; \
  }
#else
#define CNF_DEFINE_SET_TRACE(name)
#endif

}  // namespace host
}  // namespace cnf
