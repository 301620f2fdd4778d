// C ABI of libconfild_cnf.so (see include/confild_cnf.h): argument checks, kernel selection, launches.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/confild_cnf.h"
#include "layout.cuh"
#include "pack.cuh"
#include "simt.cuh"
#include "tc_kernels.cuh"
#include "tc2_kernels.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CNF_CUDA(expr)                                                                             \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) return fail(CNF_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

int check_dims(const cnf_dims* d) {
  if (!d) return fail(CNF_ERR_INVALID_ARGUMENT, "dims is NULL");
  if (d->cin < 1 || d->L < 1 || d->H < 1 || d->nl < 0 || d->cout < 1)
    return fail(CNF_ERR_INVALID_ARGUMENT, "non-positive dimension (cin=%d L=%d H=%d nl=%d cout=%d)", d->cin, d->L,
                d->H, d->nl, d->cout);
  return CNF_OK;
}

bool tc_ok(const cnf_dims& d) {
  return cnf::tc_shape_ok(d.H) && d.nl >= 1 && d.cin <= 4 && d.cout <= 4;
}

struct DeviceInfo {
  int sms = 0;
  int max_smem_optin = 0;
};
int device_info(DeviceInfo* info) {
  int dev = 0;
  CNF_CUDA(cudaGetDevice(&dev));
  CNF_CUDA(cudaDeviceGetAttribute(&info->sms, cudaDevAttrMultiProcessorCount, dev));
  CNF_CUDA(cudaDeviceGetAttribute(&info->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return CNF_OK;
}

size_t simt_smem_bytes(const cnf_dims& d) {
  const int extra = d.cin > d.cout ? d.cin : d.cout;
  return ((size_t)2 * cnf::kSimtTM * (d.H + 1) + (size_t)cnf::kSimtTM * extra) * sizeof(float);
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// Packed tiles (rows = consecutive (frame, point) pairs, a tile may span frames) when frame-aligned 128-point tiles
// would waste >= 1/6 of their rows on padding: P < 128, or a ragged P of a few hundred.  CNF_TC_PACKED=0/1 forces it.
int use_packed(int64_t P) {
  const int forced = env_int("CNF_TC_PACKED", -1);
  if (forced == 0 || forced == 1) return forced;
  const int64_t padded = (P + cnf::kTileM - 1) / cnf::kTileM * cnf::kTileM;
  return padded * 5 >= P * 6 ? 1 : 0;
}

// Launch plan for a tensor-core kernel: ring depth, shared memory, CTAs per SM, grid.
struct TcPlan {
  int stages = 0;
  size_t smem = 0;
  int ctas_per_sm = 1;
  int64_t grid = 0;
  unsigned tmem_cols = 0;
};

template <int H, int PREC>
int make_tc_plan(const DeviceInfo& di, int64_t tiles, TcPlan* plan) {
  using C = cnf::TcCfg<H, PREC>;
  // One CTA per SM (16 activation warps + issuer + producer fill the register file); the weight ring takes whatever
  // shared memory the A operand leaves: 2 stages at H=384 bf16x3, 6 at H=256 bf16x3, 12 otherwise.
  const size_t fixed = cnf::tc_smem_bytes<H, PREC>(0);
  if ((size_t)di.max_smem_optin <= fixed + 2 * cnf::kStageBytes)
    return fail(CNF_ERR_UNSUPPORTED, "H=%d precision=%d does not fit in shared memory", H, PREC);
  int stages = (int)(((size_t)di.max_smem_optin - fixed) / cnf::kStageBytes);
  if (stages > cnf::kTcMaxStages) stages = cnf::kTcMaxStages;
  const int forced = env_int("CNF_TC_STAGES", 0);
  if (forced >= 2 && forced <= stages) stages = forced;
  stages -= stages % C::kNBlocks;  // the MMA warp consumes the ring in groups of kNBlocks adjacent slots
  if (stages < C::kNBlocks) return fail(CNF_ERR_UNSUPPORTED, "weight ring too small for H=%d", H);
  plan->stages = stages;
  plan->ctas_per_sm = 1;
  plan->smem = cnf::tc_smem_bytes<H, PREC>(stages);
  plan->tmem_cols = C::kTmemCols;
  plan->grid = tiles < di.sms ? tiles : di.sms;
  return CNF_OK;
}

template <int H, int PREC, bool STASH, bool REDUCE>
int launch_tc_forward(const cnf_dims& d, const uint8_t* packed, const float* coords, int64_t cfs, const float* shift,
                      cnf::OutTargets out, void* stash, int64_t T, int64_t P, cudaStream_t st) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  const int pack_rows = use_packed(P);
  const int64_t tiles = cnf::tc_num_tiles(T, P, pack_rows);
  TcPlan plan;
  if (int rc = make_tc_plan<H, PREC>(di, tiles, &plan)) return rc;
  // frame-aligned tiles of the block-pipelined kernels stage the layer's FiLM shifts in shared memory
  constexpr bool kCanStage = cnf::TcCfg<H, PREC>::kBlockPipe;
  auto kern = (kCanStage && !pack_rows) ? cnf::tc_forward_kernel<H, PREC, STASH, REDUCE, kCanStage>
                                        : cnf::tc_forward_kernel<H, PREC, STASH, REDUCE, false>;
  CNF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
  kern<<<(unsigned)plan.grid, cnf::kTcThreads, plan.smem, st>>>(d, packed, coords, cfs, shift, out,
                                                                reinterpret_cast<__half*>(stash), T, P, plan.stages, pack_rows);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}

template <int H, int PREC>
int dispatch_tc_forward(const cnf_dims& d, const uint8_t* packed, const float* coords, int64_t cfs,
                        const float* shift, cnf::OutTargets out, void* stash, int64_t T, int64_t P, cudaStream_t st) {
  const bool reduce = env_int("CNF_TC_REDUCE", 0) != 0;
  if (stash) {
    return reduce ? launch_tc_forward<H, PREC, true, true>(d, packed, coords, cfs, shift, out, stash, T, P, st)
                  : launch_tc_forward<H, PREC, true, false>(d, packed, coords, cfs, shift, out, stash, T, P, st);
  }
  return reduce ? launch_tc_forward<H, PREC, false, true>(d, packed, coords, cfs, shift, out, stash, T, P, st)
                : launch_tc_forward<H, PREC, false, false>(d, packed, coords, cfs, shift, out, stash, T, P, st);
}

template <int PREC>
int dispatch_tc_forward_h(const cnf_dims& d, const uint8_t* packed, const float* coords, int64_t cfs,
                          const float* shift, cnf::OutTargets out, void* stash, int64_t T, int64_t P, cudaStream_t st) {
  switch (d.H) {
    case 128: return dispatch_tc_forward<128, PREC>(d, packed, coords, cfs, shift, out, stash, T, P, st);
    case 256: return dispatch_tc_forward<256, PREC>(d, packed, coords, cfs, shift, out, stash, T, P, st);
    case 384: return dispatch_tc_forward<384, PREC>(d, packed, coords, cfs, shift, out, stash, T, P, st);
  }
  return fail(CNF_ERR_UNSUPPORTED, "no tensor-core kernel for H=%d", d.H);
}

// H = 128 fast path (activations in TMEM, two tiles in flight, one CTA per SM).
bool use_tc2(const cnf_dims& d) { return d.H == cnf::kTc2H && env_int("CNF_TC2", 1) != 0; }

int make_tc2_plan(const DeviceInfo& di, int64_t tiles, TcPlan* plan) {
  const size_t fixed = cnf::tc2_smem_bytes(0);
  int stages = (int)(((size_t)di.max_smem_optin - fixed) / cnf::kStageBytes);
  if (stages > cnf::kTcMaxStages) stages = cnf::kTcMaxStages;
  const int forced = env_int("CNF_TC_STAGES", 0);
  if (forced >= 4 && forced <= stages) stages = forced;
  if (stages < 6) return fail(CNF_ERR_UNSUPPORTED, "not enough shared memory for the weight ring");
  plan->stages = stages;
  plan->ctas_per_sm = 1;
  plan->smem = cnf::tc2_smem_bytes(stages);
  plan->tmem_cols = 512;
  const int64_t pairs = (tiles + 1) / 2;
  plan->grid = pairs < di.sms ? pairs : di.sms;
  return CNF_OK;
}

template <int PREC, bool STASH, bool PACKED>
int launch_tc2_forward(const cnf_dims& d, const uint8_t* packed, const float* coords, int64_t cfs, const float* shift,
                       cnf::OutTargets out, void* stash, int64_t T, int64_t P, cudaStream_t st) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  const int64_t tiles = cnf::tc_num_tiles(T, P, PACKED ? 1 : 0);
  TcPlan plan;
  if (int rc = make_tc2_plan(di, tiles, &plan)) return rc;
  auto kern = cnf::tc2_forward_kernel<PREC, STASH, PACKED>;
  CNF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
  kern<<<(unsigned)plan.grid, cnf::kTc2Threads, plan.smem, st>>>(d, packed, coords, cfs, shift, out,
                                                                 reinterpret_cast<__half*>(stash), T, P, plan.stages);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}

template <int PREC>
int dispatch_tc2_forward(const cnf_dims& d, const uint8_t* packed, const float* coords, int64_t cfs,
                         const float* shift, cnf::OutTargets out, void* stash, int64_t T, int64_t P, cudaStream_t st) {
  const bool pk = use_packed(P) != 0;
  if (stash)
    return pk ? launch_tc2_forward<PREC, true, true>(d, packed, coords, cfs, shift, out, stash, T, P, st)
              : launch_tc2_forward<PREC, true, false>(d, packed, coords, cfs, shift, out, stash, T, P, st);
  return pk ? launch_tc2_forward<PREC, false, true>(d, packed, coords, cfs, shift, out, stash, T, P, st)
            : launch_tc2_forward<PREC, false, false>(d, packed, coords, cfs, shift, out, stash, T, P, st);
}

template <bool PACKED>
int launch_tc2_backward_t(const cnf_dims& d, const uint8_t* packed, const float* gout, const void* stash, float* gshift,
                          int64_t T, int64_t P, cudaStream_t st) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  const int64_t tiles = cnf::tc_num_tiles(T, P, PACKED ? 1 : 0);
  TcPlan plan;
  if (int rc = make_tc2_plan(di, tiles, &plan)) return rc;
  auto kern = cnf::tc2_backward_kernel<PACKED>;
  CNF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
  kern<<<(unsigned)plan.grid, cnf::kTc2BwdThreads, plan.smem, st>>>(d, packed, gout, reinterpret_cast<const __half*>(stash),
                                                                 gshift, T, P, plan.stages);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}

int launch_tc2_backward(const cnf_dims& d, const uint8_t* packed, const float* gout, const void* stash, float* gshift,
                        int64_t T, int64_t P, cudaStream_t st) {
  return use_packed(P) ? launch_tc2_backward_t<true>(d, packed, gout, stash, gshift, T, P, st)
                       : launch_tc2_backward_t<false>(d, packed, gout, stash, gshift, T, P, st);
}

template <int H>
int launch_tc_backward(const cnf_dims& d, const uint8_t* packed, const float* gout, const void* stash, float* gshift,
                       int64_t T, int64_t P, cudaStream_t st) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  const int pack_rows = use_packed(P);
  const int64_t tiles = cnf::tc_num_tiles(T, P, pack_rows);
  TcPlan plan;
  if (int rc = make_tc_plan<H, CNF_PREC_BF16X3>(di, tiles, &plan)) return rc;
  auto kern = cnf::tc_backward_kernel<H>;
  CNF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
  kern<<<(unsigned)plan.grid, cnf::kTcThreads, plan.smem, st>>>(d, packed, gout, reinterpret_cast<const __half*>(stash),
                                                                gshift, T, P, plan.stages, pack_rows);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}

int64_t simt_grid(int64_t tiles, int sms) {
  const int64_t cap = (int64_t)sms * 4;
  return tiles < cap ? tiles : cap;
}

}  // namespace

#ifdef CNF_TRACE
extern "C" int cnf_debug_set_trace(void* d_buf) {
  unsigned long long* p = static_cast<unsigned long long*>(d_buf);
  return cudaMemcpyToSymbol(cnf::g_trace, &p, sizeof(p)) == cudaSuccess ? 0 : 3;
}
#endif

extern "C" {

int cnf_abi_version(void) { return CNF_ABI_VERSION; }

const char* cnf_last_error(void) { return g_err; }

int cnf_tc_supported(const cnf_dims* dims) {
  if (check_dims(dims)) return 0;
  return tc_ok(*dims) ? 1 : 0;
}

int cnf_param_count(const cnf_dims* dims, size_t* count) {
  if (int rc = check_dims(dims)) return rc;
  if (!count) return fail(CNF_ERR_INVALID_ARGUMENT, "count is NULL");
  *count = cnf::make_param_offsets(*dims).total;
  return CNF_OK;
}

int cnf_packed_bytes(const cnf_dims* dims, size_t* bytes) {
  if (int rc = check_dims(dims)) return rc;
  if (!bytes) return fail(CNF_ERR_INVALID_ARGUMENT, "bytes is NULL");
  *bytes = cnf::make_layout(*dims).total;
  return CNF_OK;
}

int cnf_pack_weights(const cnf_dims* dims, const float* d_params_flat, float w0, void* d_packed, size_t packed_bytes,
                     void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!d_params_flat || !d_packed) return fail(CNF_ERR_INVALID_ARGUMENT, "NULL device pointer");
  const cnf::PackedLayout lay = cnf::make_layout(*dims);
  if (packed_bytes < lay.total)
    return fail(CNF_ERR_BUFFER_TOO_SMALL, "packed buffer has %zu bytes, need %zu", packed_bytes, lay.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* packed = static_cast<uint8_t*>(d_packed);
  cnf::pack_fp32_kernel<<<296, 256, 0, st>>>(*dims, d_params_flat, w0, packed);
  CNF_CUDA(cudaGetLastError());
  if (cnf::tc_shape_ok(dims->H) && dims->nl >= 1) {
    for (int mode = 0; mode < 3; ++mode) {
      cnf::pack_tc_kernel<<<592, 256, 0, st>>>(*dims, d_params_flat, w0, packed, mode);
      CNF_CUDA(cudaGetLastError());
    }
  }
  return CNF_OK;
}

int cnf_film_shift(const cnf_dims* dims, const void* d_packed, const float* d_latents, int64_t T, float* d_shift,
                   void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!d_packed || !d_latents || !d_shift) return fail(CNF_ERR_INVALID_ARGUMENT, "NULL device pointer");
  if (T < 1) return fail(CNF_ERR_INVALID_ARGUMENT, "T=%lld", (long long)T);
  const cnf::PackedLayout lay = cnf::make_layout(*dims);
  const uint8_t* packed = static_cast<const uint8_t*>(d_packed);
  const int N = (dims->nl + 1) * dims->H;
  dim3 grid((N + 63) / 64, (unsigned)((T + 63) / 64));
  cnf::simt_gemm_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_latents, reinterpret_cast<const float*>(packed + lay.v_cat),
      reinterpret_cast<const float*>(packed + lay.b_shift), d_shift, T, N, dims->L);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}

int cnf_film_shift_backward(const cnf_dims* dims, const void* d_packed, const float* d_gshift, int64_t T,
                            float* d_glatents, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!d_packed || !d_gshift || !d_glatents) return fail(CNF_ERR_INVALID_ARGUMENT, "NULL device pointer");
  if (T < 1) return fail(CNF_ERR_INVALID_ARGUMENT, "T=%lld", (long long)T);
  const cnf::PackedLayout lay = cnf::make_layout(*dims);
  const uint8_t* packed = static_cast<const uint8_t*>(d_packed);
  const int K = (dims->nl + 1) * dims->H;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CNF_CUDA(cudaMemsetAsync(d_glatents, 0, (size_t)T * dims->L * sizeof(float), st));
  if (T > 2147483647LL) return fail(CNF_ERR_UNSUPPORTED, "T=%lld exceeds the grid limit", (long long)T);
  dim3 grid((unsigned)T, (unsigned)((K + cnf::kShiftBwdChunk - 1) / cnf::kShiftBwdChunk));
  cnf::film_shift_backward_kernel<<<grid, 128, 0, st>>>(d_gshift, reinterpret_cast<const float*>(packed + lay.v_cat),
                                                       d_glatents, K, dims->L);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}

int cnf_stash_bytes(const cnf_dims* dims, int precision, int64_t T, int64_t P, size_t* bytes) {
  if (int rc = check_dims(dims)) return rc;
  if (!bytes) return fail(CNF_ERR_INVALID_ARGUMENT, "bytes is NULL");
  if (T < 1 || P < 1) return fail(CNF_ERR_INVALID_ARGUMENT, "T=%lld P=%lld", (long long)T, (long long)P);
  size_t esize;
  switch (precision) {
    case CNF_PREC_FP32: esize = 4; break;
    case CNF_PREC_BF16X3:
    case CNF_PREC_FP16: esize = 2; break;
    default: return fail(CNF_ERR_INVALID_ARGUMENT, "unknown precision %d", precision);
  }
  // fp32 path: [t][p][layer][column]; tensor-core paths: tile-major with rows padded to whole 128-point tiles
  const size_t rows = precision == CNF_PREC_FP32 ? (size_t)P : (size_t)((P + cnf::kTileM - 1) / cnf::kTileM) * cnf::kTileM;
  *bytes = (size_t)T * rows * (size_t)(dims->nl + 1) * (size_t)dims->H * esize;
  return CNF_OK;
}

static int forward_impl(const cnf_dims* dims, const void* d_packed, int precision, const float* d_coords,
                        int64_t coord_frame_stride, const float* d_shift, cnf::OutTargets outs, int64_t T, int64_t P,
                        void* d_stash, size_t stash_bytes, void* stream) {
  float* d_out = outs.ptr[0];
  if (int rc = check_dims(dims)) return rc;
  if (!d_packed || !d_coords || !d_shift || !d_out) return fail(CNF_ERR_INVALID_ARGUMENT, "NULL device pointer");
  if (T < 1 || P < 1) return fail(CNF_ERR_INVALID_ARGUMENT, "T=%lld P=%lld", (long long)T, (long long)P);
  if (coord_frame_stride < 0) return fail(CNF_ERR_INVALID_ARGUMENT, "negative coord_frame_stride");
  if (d_stash) {
    size_t need = 0;
    if (int rc = cnf_stash_bytes(dims, precision, T, P, &need)) return rc;
    if (stash_bytes < need) return fail(CNF_ERR_BUFFER_TOO_SMALL, "stash has %zu bytes, need %zu", stash_bytes, need);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t* packed = static_cast<const uint8_t*>(d_packed);
  if (precision == CNF_PREC_FP32) {
    if (outs.n != 1) return fail(CNF_ERR_UNSUPPORTED, "the fused gather needs a tensor-core precision");
    DeviceInfo di;
    if (int rc = device_info(&di)) return rc;
    const size_t smem = simt_smem_bytes(*dims);
    if (smem > (size_t)di.max_smem_optin)
      return fail(CNF_ERR_UNSUPPORTED, "H=%d needs %zu bytes of shared memory (> %d)", dims->H, smem, di.max_smem_optin);
    const int64_t tiles = T * ((P + cnf::kSimtTM - 1) / cnf::kSimtTM);
    const unsigned grid = (unsigned)simt_grid(tiles, di.sms);
    if (d_stash) {
      CNF_CUDA(cudaFuncSetAttribute(cnf::simt_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cnf::simt_forward_kernel<true><<<grid, 256, smem, st>>>(*dims, packed, d_coords, coord_frame_stride, d_shift,
                                                             d_out, static_cast<float*>(d_stash), T, P);
    } else {
      CNF_CUDA(cudaFuncSetAttribute(cnf::simt_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cnf::simt_forward_kernel<false><<<grid, 256, smem, st>>>(*dims, packed, d_coords, coord_frame_stride, d_shift,
                                                              d_out, nullptr, T, P);
    }
    CNF_CUDA(cudaGetLastError());
    return CNF_OK;
  }
  if (precision != CNF_PREC_BF16X3 && precision != CNF_PREC_FP16)
    return fail(CNF_ERR_INVALID_ARGUMENT, "unknown precision %d", precision);
  if (!tc_ok(*dims))
    return fail(CNF_ERR_UNSUPPORTED,
                "tensor-core path needs H in {128,256,384}, nl>=1, cin<=4, cout<=4 (got H=%d nl=%d cin=%d cout=%d); "
                "use CNF_PREC_FP32",
                dims->H, dims->nl, dims->cin, dims->cout);
  if (use_tc2(*dims)) {
    return precision == CNF_PREC_BF16X3
               ? dispatch_tc2_forward<CNF_PREC_BF16X3>(*dims, packed, d_coords, coord_frame_stride, d_shift, outs,
                                                       d_stash, T, P, st)
               : dispatch_tc2_forward<CNF_PREC_FP16>(*dims, packed, d_coords, coord_frame_stride, d_shift, outs,
                                                     d_stash, T, P, st);
  }
  if (precision == CNF_PREC_BF16X3)
    return dispatch_tc_forward_h<CNF_PREC_BF16X3>(*dims, packed, d_coords, coord_frame_stride, d_shift, outs, d_stash,
                                                  T, P, st);
  return dispatch_tc_forward_h<CNF_PREC_FP16>(*dims, packed, d_coords, coord_frame_stride, d_shift, outs, d_stash, T,
                                              P, st);
}

int cnf_forward(const cnf_dims* dims, const void* d_packed, int precision, const float* d_coords,
                int64_t coord_frame_stride, const float* d_shift, float* d_out, int64_t T, int64_t P, void* d_stash,
                size_t stash_bytes, void* stream) {
  cnf::OutTargets outs{};
  outs.ptr[0] = d_out;
  outs.n = 1;
  return forward_impl(dims, d_packed, precision, d_coords, coord_frame_stride, d_shift, outs, T, P, d_stash,
                      stash_bytes, stream);
}

int cnf_forward_gather(const cnf_dims* dims, const void* d_packed, int precision, const float* d_coords,
                       int64_t coord_frame_stride, const float* d_shift, float* const* d_outs, int n_out, int64_t T,
                       int64_t P, void* stream) {
  if (!d_outs) return fail(CNF_ERR_INVALID_ARGUMENT, "d_outs is NULL");
  if (n_out < 1 || n_out > cnf::kMaxOutTargets)
    return fail(CNF_ERR_INVALID_ARGUMENT, "n_out=%d, must be in [1,%d]", n_out, cnf::kMaxOutTargets);
  cnf::OutTargets outs{};
  for (int k = 0; k < n_out; ++k) {
    if (!d_outs[k]) return fail(CNF_ERR_INVALID_ARGUMENT, "d_outs[%d] is NULL", k);
    outs.ptr[k] = d_outs[k];
  }
  outs.n = n_out;
  return forward_impl(dims, d_packed, precision, d_coords, coord_frame_stride, d_shift, outs, T, P, nullptr, 0, stream);
}

int cnf_backward(const cnf_dims* dims, const void* d_packed, int precision, const float* d_gout, const void* d_stash,
                 size_t stash_bytes, float* d_gshift, int64_t T, int64_t P, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!d_packed || !d_gout || !d_stash || !d_gshift) return fail(CNF_ERR_INVALID_ARGUMENT, "NULL device pointer");
  if (T < 1 || P < 1) return fail(CNF_ERR_INVALID_ARGUMENT, "T=%lld P=%lld", (long long)T, (long long)P);
  size_t need = 0;
  if (int rc = cnf_stash_bytes(dims, precision, T, P, &need)) return rc;
  if (stash_bytes < need) return fail(CNF_ERR_BUFFER_TOO_SMALL, "stash has %zu bytes, need %zu", stash_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t* packed = static_cast<const uint8_t*>(d_packed);
  const size_t gbytes = (size_t)T * (size_t)(dims->nl + 1) * (size_t)dims->H * sizeof(float);
  CNF_CUDA(cudaMemsetAsync(d_gshift, 0, gbytes, st));
  if (precision == CNF_PREC_FP32) {
    DeviceInfo di;
    if (int rc = device_info(&di)) return rc;
    const size_t smem = simt_smem_bytes(*dims);
    if (smem > (size_t)di.max_smem_optin)
      return fail(CNF_ERR_UNSUPPORTED, "H=%d needs %zu bytes of shared memory (> %d)", dims->H, smem, di.max_smem_optin);
    const int64_t tiles = T * ((P + cnf::kSimtTM - 1) / cnf::kSimtTM);
    CNF_CUDA(cudaFuncSetAttribute(cnf::simt_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cnf::simt_backward_kernel<<<(unsigned)simt_grid(tiles, di.sms), 256, smem, st>>>(
        *dims, packed, d_gout, static_cast<const float*>(d_stash), d_gshift, T, P);
    CNF_CUDA(cudaGetLastError());
    return CNF_OK;
  }
  if (!tc_ok(*dims))
    return fail(CNF_ERR_UNSUPPORTED, "tensor-core path unsupported for H=%d nl=%d cin=%d cout=%d; use CNF_PREC_FP32",
                dims->H, dims->nl, dims->cin, dims->cout);
  if (use_tc2(*dims)) return launch_tc2_backward(*dims, packed, d_gout, d_stash, d_gshift, T, P, st);
  switch (dims->H) {
    case 128: return launch_tc_backward<128>(*dims, packed, d_gout, d_stash, d_gshift, T, P, st);
    case 256: return launch_tc_backward<256>(*dims, packed, d_gout, d_stash, d_gshift, T, P, st);
    case 384: return launch_tc_backward<384>(*dims, packed, d_gout, d_stash, d_gshift, T, P, st);
  }
  return fail(CNF_ERR_UNSUPPORTED, "no tensor-core kernel for H=%d", dims->H);
}

int cnf_query_launch(const cnf_dims* dims, int precision, int64_t T, int64_t P, int64_t* values, int n) {
  if (int rc = check_dims(dims)) return rc;
  if (!values || n < 1) return fail(CNF_ERR_INVALID_ARGUMENT, "values is NULL or n < 1");
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  int64_t v[7] = {di.sms, 0, 0, 0, 0, 0, 0};
  if (precision == CNF_PREC_FP32) {
    const int64_t tiles = T * ((P + cnf::kSimtTM - 1) / cnf::kSimtTM);
    v[1] = simt_grid(tiles, di.sms);
    v[2] = 256;
    v[3] = (int64_t)simt_smem_bytes(*dims);
    v[4] = 0;
    v[5] = 0;
    v[6] = cnf::kSimtTM;
  } else {
    if (!tc_ok(*dims)) return fail(CNF_ERR_UNSUPPORTED, "tensor-core path unsupported for these dims");
    const int64_t tiles = cnf::tc_num_tiles(T, P, use_packed(P));
    TcPlan plan;
    int rc = CNF_ERR_UNSUPPORTED;
    const bool x3 = precision == CNF_PREC_BF16X3;
    if (use_tc2(*dims)) {
      if (int rc2 = make_tc2_plan(di, tiles, &plan)) return rc2;
      const int64_t v2[7] = {di.sms, plan.grid, cnf::kTc2Threads, (int64_t)plan.smem, 1, 512, 2 * cnf::kTileM};
      for (int i = 0; i < n && i < 7; ++i) values[i] = v2[i];
      return CNF_OK;
    }
    switch (dims->H) {
      case 128: rc = x3 ? make_tc_plan<128, CNF_PREC_BF16X3>(di, tiles, &plan) : make_tc_plan<128, CNF_PREC_FP16>(di, tiles, &plan); break;
      case 256: rc = x3 ? make_tc_plan<256, CNF_PREC_BF16X3>(di, tiles, &plan) : make_tc_plan<256, CNF_PREC_FP16>(di, tiles, &plan); break;
      case 384: rc = x3 ? make_tc_plan<384, CNF_PREC_BF16X3>(di, tiles, &plan) : make_tc_plan<384, CNF_PREC_FP16>(di, tiles, &plan); break;
    }
    if (rc) return rc;
    v[1] = plan.grid;
    v[2] = cnf::kTcThreads;
    v[3] = (int64_t)plan.smem;
    v[4] = plan.ctas_per_sm;
    v[5] = plan.tmem_cols;
    v[6] = cnf::kTileM;
  }
  for (int i = 0; i < n && i < 7; ++i) values[i] = v[i];
  return CNF_OK;
}

}  // extern "C"
