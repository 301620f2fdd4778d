// C ABI of libconfild_cnf.so (see include/confild_cnf.h): argument checks, kernel selection, launches of the
// CUDA-core kernels.  The tensor-core kernels are launched from their own translation units (host.cuh).
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/confild_cnf.h"
#include "host.cuh"
#include "layout.cuh"
#include "pack.cuh"
#include "simt.cuh"

namespace cnf {
namespace host {

namespace {
thread_local char g_err[512] = "";
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int device_info(DeviceInfo* info) {
  static DeviceInfo cache[kMaxDevices];
  static std::atomic<int> have[kMaxDevices];
  int dev = 0;
  CNF_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < kMaxDevices && have[dev].load(std::memory_order_acquire)) {
    *info = cache[dev];
    return CNF_OK;
  }
  DeviceInfo di;
  di.device = dev;
  CNF_CUDA(cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, dev));
  CNF_CUDA(cudaDeviceGetAttribute(&di.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (dev >= 0 && dev < kMaxDevices) {
    cache[dev] = di;
    have[dev].store(1, std::memory_order_release);
  }
  *info = di;
  return CNF_OK;
}

namespace {
Knobs& mutable_knobs() {
  static Knobs k;
  static std::once_flag once;
  std::call_once(once, [] {  // the environment gives the initial values, once per process
    auto env_int = [](const char* name, int dflt) {
      const char* v = getenv(name);
      return v ? atoi(v) : dflt;
    };
    k.tc2 = env_int("CNF_TC2", 1);
    k.stages = env_int("CNF_TC_STAGES", 0);
    k.packed = env_int("CNF_TC_PACKED", -1);
    k.cluster = env_int("CNF_TC_CLUSTER", 1);
    k.gn_cluster = env_int("CNF_GN_CLUSTER", 256);
  });
  return k;
}
}  // namespace
const Knobs& knobs() { return mutable_knobs(); }
int set_knob(const char* name, int value) {
  Knobs& k = mutable_knobs();
  if (!strcmp(name, "CNF_TC2")) k.tc2 = value;
  else if (!strcmp(name, "CNF_TC_STAGES")) k.stages = value;
  else if (!strcmp(name, "CNF_TC_PACKED")) k.packed = value;
  else if (!strcmp(name, "CNF_TC_CLUSTER")) k.cluster = value;
  else if (!strcmp(name, "CNF_GN_CLUSTER")) k.gn_cluster = value;
  else return fail(CNF_ERR_INVALID_ARGUMENT, "unknown debug knob %s", name);
  return CNF_OK;
}

}  // namespace host
}  // namespace cnf

namespace {

using cnf::host::BwdArgs;
using cnf::host::DeviceInfo;
using cnf::host::device_info;
using cnf::host::fail;
using cnf::host::FwdArgs;
using cnf::host::LaunchInfo;

int check_dims(const cnf_dims* d) {
  if (!d) return fail(CNF_ERR_INVALID_ARGUMENT, "dims is NULL");
  if (d->cin < 1 || d->L < 1 || d->H < 1 || d->nl < 0 || d->cout < 1)
    return fail(CNF_ERR_INVALID_ARGUMENT, "non-positive dimension (cin=%d L=%d H=%d nl=%d cout=%d)", d->cin, d->L,
                d->H, d->nl, d->cout);
  return CNF_OK;
}

bool tc_ok(const cnf_dims& d) { return cnf::tc_shape_ok(d.H) && d.nl >= 1 && d.cin <= 4 && d.cout <= 4; }

bool is_tc_precision(int precision) {
  return precision == CNF_PREC_BF16X3 || precision == CNF_PREC_FP16 || precision == CNF_PREC_F16F8;
}

// H = 128 fast path (activations in TMEM, two tiles in flight, one CTA per SM).
bool use_tc2(const cnf_dims& d) { return d.H == 128 && cnf::host::knobs().tc2 != 0; }

size_t simt_smem_bytes(const cnf_dims& d) {
  const int extra = d.cin > d.cout ? d.cin : d.cout;
  return ((size_t)2 * cnf::kSimtTM * (d.H + 1) + (size_t)cnf::kSimtTM * extra) * sizeof(float);
}

int64_t simt_grid(int64_t tiles, int sms) {
  const int64_t cap = (int64_t)sms * 4;
  return tiles < cap ? tiles : cap;
}

int tc_forward_dispatch(int precision, const FwdArgs& a) {
  if (use_tc2(a.d)) {
    switch (precision) {
      case CNF_PREC_BF16X3: return cnf::host::tc2_forward_bf16x3(a);
      case CNF_PREC_F16F8: return cnf::host::tc2_forward_f16f8(a);
      default: return cnf::host::tc2_forward_fp16(a);
    }
  }
  switch (precision) {
    case CNF_PREC_BF16X3: return cnf::host::tc_forward_bf16x3(a);
    case CNF_PREC_F16F8: return cnf::host::tc_forward_f16f8(a);
    default: return cnf::host::tc_forward_fp16(a);
  }
}

template <bool STASH>
int launch_simt_forward(const FwdArgs& a, const DeviceInfo& di, size_t smem) {
  static std::atomic<size_t> smem_set[cnf::host::kMaxDevices];
  const int64_t tiles = a.T * ((a.P + cnf::kSimtTM - 1) / cnf::kSimtTM);
  if (int rc = cnf::host::ensure_smem(cnf::simt_forward_kernel<STASH>, smem, di.device, smem_set)) return rc;
  cnf::simt_forward_kernel<STASH><<<(unsigned)simt_grid(tiles, di.sms), 256, smem, a.stream>>>(
      a.d, a.packed, a.coords, a.coord_frame_stride, a.shift, a.outs.ptr[0], static_cast<float*>(a.stash), a.T, a.P);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}

int forward_impl(const cnf_dims* dims, const void* d_packed, int precision, const float* d_coords,
                 int64_t coord_frame_stride, const float* d_shift, cnf::OutTargets outs, int64_t T, int64_t P,
                 void* d_stash, size_t stash_bytes, const cnf::LossArgs* loss, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  const bool with_loss = loss != nullptr && loss->y_meas != nullptr;
  if (!d_packed || !d_coords || !d_shift) return fail(CNF_ERR_INVALID_ARGUMENT, "NULL device pointer");
  if (!outs.ptr[0] && !(with_loss && is_tc_precision(precision)))
    return fail(CNF_ERR_INVALID_ARGUMENT, "NULL output pointer");
  if (T < 1 || P < 1) return fail(CNF_ERR_INVALID_ARGUMENT, "T=%lld P=%lld", (long long)T, (long long)P);
  if (coord_frame_stride < 0) return fail(CNF_ERR_INVALID_ARGUMENT, "negative coord_frame_stride");
  if (d_stash) {
    size_t need = 0;
    if (int rc = cnf_stash_bytes(dims, precision, T, P, &need)) return rc;
    if (stash_bytes < need) return fail(CNF_ERR_BUFFER_TOO_SMALL, "stash has %zu bytes, need %zu", stash_bytes, need);
  }
  FwdArgs a{};
  a.d = *dims;
  a.packed = static_cast<const uint8_t*>(d_packed);
  a.coords = d_coords;
  a.coord_frame_stride = coord_frame_stride;
  a.shift = d_shift;
  a.outs = outs;
  a.stash = d_stash;
  if (with_loss) a.loss = *loss;
  a.T = T;
  a.P = P;
  a.stream = static_cast<cudaStream_t>(stream);
  a.query = nullptr;
  if (precision == CNF_PREC_FP32) {
    if (outs.n != 1) return fail(CNF_ERR_UNSUPPORTED, "the fused gather needs a tensor-core precision");
    DeviceInfo di;
    if (int rc = device_info(&di)) return rc;
    const size_t smem = simt_smem_bytes(*dims);
    if (smem > (size_t)di.max_smem_optin)
      return fail(CNF_ERR_UNSUPPORTED, "H=%d needs %zu bytes of shared memory (> %d)", dims->H, smem, di.max_smem_optin);
    if (int rc = d_stash ? launch_simt_forward<true>(a, di, smem) : launch_simt_forward<false>(a, di, smem)) return rc;
    if (with_loss) {  // the fp32 path evaluates the loss in a separate pass over the decoded field
      const int64_t rows = T * P;
      int64_t blocks = (rows + 255) / 256;
      if (blocks > 1024) blocks = 1024;
      cnf::loss_rows_kernel<<<(unsigned)blocks, 256, 0, a.stream>>>(a.loss, outs.ptr[0], T, P, dims->cout);
      CNF_CUDA(cudaGetLastError());
    }
    return CNF_OK;
  }
  if (!is_tc_precision(precision)) return fail(CNF_ERR_INVALID_ARGUMENT, "unknown precision %d", precision);
  if (!tc_ok(*dims))
    return fail(CNF_ERR_UNSUPPORTED,
                "tensor-core path needs H in {128,256,384}, nl>=1, cin<=4, cout<=4 (got H=%d nl=%d cin=%d cout=%d); "
                "use CNF_PREC_FP32",
                dims->H, dims->nl, dims->cin, dims->cout);
  return tc_forward_dispatch(precision, a);
}

}  // namespace

#ifdef CNF_TRACE
extern "C" int cnf_debug_set_trace(void* d_buf) {
  unsigned long long* p = static_cast<unsigned long long*>(d_buf);
  int rc = 0;
  rc |= cnf::host::set_trace_tc2_fwd_bf16x3(p);
  rc |= cnf::host::set_trace_tc2_fwd_fp16(p);
  rc |= cnf::host::set_trace_tc2_fwd_f16f8(p);
  rc |= cnf::host::set_trace_tc_fwd_f16f8(p);
  rc |= cnf::host::set_trace_tc2_bwd(p);
  rc |= cnf::host::set_trace_tc_fwd_bf16x3(p);
  rc |= cnf::host::set_trace_tc_fwd_fp16(p);
  rc |= cnf::host::set_trace_tc_bwd(p);
  return rc;
}
#endif

extern "C" {

int cnf_abi_version(void) { return CNF_ABI_VERSION; }

int cnf_set_debug_knob(const char* name, int value) {
  if (!name) return fail(CNF_ERR_INVALID_ARGUMENT, "name is NULL");
  return cnf::host::set_knob(name, value);
}

const char* cnf_last_error(void) { return cnf::host::g_err; }

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
    cnf::pack_scale_kernel<<<dims->nl, 256, 0, st>>>(*dims, d_params_flat, w0, packed);
    CNF_CUDA(cudaGetLastError());
    cnf::pack_tc_f8_kernel<<<592, 256, 0, st>>>(*dims, d_params_flat, w0, packed);
    CNF_CUDA(cudaGetLastError());
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
  if (grid.y > 65535) return fail(CNF_ERR_UNSUPPORTED, "T=%lld exceeds the grid limit", (long long)T);
  cnf::simt_gemm_kernel<true, false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_latents, reinterpret_cast<const float*>(packed + lay.v_cat),
      reinterpret_cast<const float*>(packed + lay.b_shift), d_shift, T, N, dims->L, dims->L, nullptr);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}

int cnf_film_shift_backward_scaled(const cnf_dims* dims, const void* d_packed, const float* d_gshift, int64_t T,
                                   const float* d_scale, float* d_glatents, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!d_packed || !d_gshift || !d_glatents) return fail(CNF_ERR_INVALID_ARGUMENT, "NULL device pointer");
  if (T < 1) return fail(CNF_ERR_INVALID_ARGUMENT, "T=%lld", (long long)T);
  const cnf::PackedLayout lay = cnf::make_layout(*dims);
  const uint8_t* packed = static_cast<const uint8_t*>(d_packed);
  const int K = (dims->nl + 1) * dims->H;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CNF_CUDA(cudaMemsetAsync(d_glatents, 0, (size_t)T * dims->L * sizeof(float), st));
  // glat (T, L) = gshift (T, K) . V (K, L): few output tiles, long reduction -> split K until ~2 waves of CTAs exist
  const int64_t tiles = ((T + 63) / 64) * ((dims->L + 63) / 64);
  if ((T + 63) / 64 > 65535) return fail(CNF_ERR_UNSUPPORTED, "T=%lld exceeds the grid limit", (long long)T);
  cnf::host::DeviceInfo di;
  if (int rc = cnf::host::device_info(&di)) return rc;
  int64_t splits = (2 * (int64_t)di.sms + tiles - 1) / tiles;
  const int64_t max_splits = (K + 63) / 64;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int kc = (int)(((K + splits - 1) / splits + 15) / 16 * 16);
  splits = (K + kc - 1) / kc;
  dim3 grid((unsigned)((dims->L + 63) / 64), (unsigned)((T + 63) / 64), (unsigned)splits);
  cnf::simt_gemm_kernel<false, true><<<grid, 256, 0, st>>>(d_gshift, reinterpret_cast<const float*>(packed + lay.v_cat),
                                                          nullptr, d_glatents, T, dims->L, K, kc, d_scale);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}

int cnf_film_shift_backward(const cnf_dims* dims, const void* d_packed, const float* d_gshift, int64_t T,
                            float* d_glatents, void* stream) {
  return cnf_film_shift_backward_scaled(dims, d_packed, d_gshift, T, nullptr, d_glatents, stream);
}

int cnf_stash_bytes(const cnf_dims* dims, int precision, int64_t T, int64_t P, size_t* bytes) {
  if (int rc = check_dims(dims)) return rc;
  if (!bytes) return fail(CNF_ERR_INVALID_ARGUMENT, "bytes is NULL");
  if (T < 1 || P < 1) return fail(CNF_ERR_INVALID_ARGUMENT, "T=%lld P=%lld", (long long)T, (long long)P);
  size_t esize;
  if (precision == CNF_PREC_FP32) esize = 4;
  else if (is_tc_precision(precision)) esize = 2;
  else return fail(CNF_ERR_INVALID_ARGUMENT, "unknown precision %d", precision);
  // fp32 path: [t][p][layer][column]; tensor-core paths: tile-major, whole 128-row tiles (frame-aligned or packed)
  size_t rows;
  if (precision == CNF_PREC_FP32) rows = (size_t)T * (size_t)P;
  else rows = (size_t)T * (size_t)((P + cnf::kTileM - 1) / cnf::kTileM) * cnf::kTileM;  // >= the packed-tile count too
  *bytes = rows * (size_t)(dims->nl + 1) * (size_t)dims->H * esize;
  return CNF_OK;
}

int cnf_forward(const cnf_dims* dims, const void* d_packed, int precision, const float* d_coords,
                int64_t coord_frame_stride, const float* d_shift, float* d_out, int64_t T, int64_t P, void* d_stash,
                size_t stash_bytes, void* stream) {
  cnf::OutTargets outs{};
  outs.ptr[0] = d_out;
  outs.n = 1;
  outs.vec_ok = 1;
  if (!d_out) return fail(CNF_ERR_INVALID_ARGUMENT, "NULL device pointer");
  return forward_impl(dims, d_packed, precision, d_coords, coord_frame_stride, d_shift, outs, T, P, d_stash,
                      stash_bytes, nullptr, stream);
}

int cnf_forward_loss(const cnf_dims* dims, const void* d_packed, int precision, const float* d_coords,
                     int64_t coord_frame_stride, const float* d_shift, float* d_out, int64_t T, int64_t P,
                     void* d_stash, size_t stash_bytes, const cnf_sensor_loss* loss, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!loss) return fail(CNF_ERR_INVALID_ARGUMENT, "loss is NULL");
  if (!loss->d_y_meas || !loss->d_gy || !loss->d_partials || !loss->d_norm)
    return fail(CNF_ERR_INVALID_ARGUMENT, "NULL pointer in cnf_sensor_loss");
  if (loss->d_mask && (loss->mask_kind < 1 || loss->mask_kind > 3))
    return fail(CNF_ERR_INVALID_ARGUMENT, "mask_kind=%d, must be 1, 2 or 3", loss->mask_kind);
  if (precision == CNF_PREC_FP32 && !d_out)
    return fail(CNF_ERR_INVALID_ARGUMENT, "CNF_PREC_FP32 evaluates the loss from the decoded field: d_out is required");
  cnf::LossArgs la{};
  la.y_meas = loss->d_y_meas;
  la.mask = loss->d_mask;
  la.mask_kind = loss->d_mask ? loss->mask_kind : 0;
  for (int o = 0; o < 4; ++o) {
    la.ya[o] = loss->y_scale[o];
    la.yb[o] = loss->y_offset[o];
  }
  la.gy = loss->d_gy;
  la.partials = loss->d_partials;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CNF_CUDA(cudaMemsetAsync(loss->d_partials, 0, CNF_LOSS_PARTIALS * sizeof(float), st));
  cnf::OutTargets outs{};
  outs.ptr[0] = d_out;
  outs.n = 1;
  outs.vec_ok = 1;
  if (int rc = forward_impl(dims, d_packed, precision, d_coords, coord_frame_stride, d_shift, outs, T, P, d_stash,
                            stash_bytes, &la, stream))
    return rc;
  cnf::loss_finalize_kernel<<<1, 1024, 0, st>>>(loss->d_partials, CNF_LOSS_PARTIALS, loss->d_norm, loss->d_extra_sq);
  CNF_CUDA(cudaGetLastError());
  return CNF_OK;
}

int cnf_forward_gather(const cnf_dims* dims, const void* d_packed, int precision, const float* d_coords,
                       int64_t coord_frame_stride, const float* d_shift, float* const* d_outs, int n_out, int64_t T,
                       int64_t P, void* stream) {
  if (!d_outs) return fail(CNF_ERR_INVALID_ARGUMENT, "d_outs is NULL");
  if (n_out < 1 || n_out > cnf::kMaxOutTargets)
    return fail(CNF_ERR_INVALID_ARGUMENT, "n_out=%d, must be in [1,%d]", n_out, cnf::kMaxOutTargets);
  cnf::OutTargets outs{};
  outs.vec_ok = 1;
  for (int k = 0; k < n_out; ++k) {
    if (!d_outs[k]) return fail(CNF_ERR_INVALID_ARGUMENT, "d_outs[%d] is NULL", k);
    if (reinterpret_cast<uintptr_t>(d_outs[k]) & 3u)
      return fail(CNF_ERR_INVALID_ARGUMENT, "d_outs[%d] is not 4-byte aligned", k);
    outs.ptr[k] = d_outs[k];
    // 16-byte vector stores need every target in the same 16-byte phase as target 0; otherwise scalar stores
    if ((reinterpret_cast<uintptr_t>(d_outs[k]) & 15u) != (reinterpret_cast<uintptr_t>(d_outs[0]) & 15u)) outs.vec_ok = 0;
  }
  outs.n = n_out;
  return forward_impl(dims, d_packed, precision, d_coords, coord_frame_stride, d_shift, outs, T, P, nullptr, 0, nullptr,
                      stream);
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
    static std::atomic<size_t> smem_set[cnf::host::kMaxDevices];
    DeviceInfo di;
    if (int rc = device_info(&di)) return rc;
    const size_t smem = simt_smem_bytes(*dims);
    if (smem > (size_t)di.max_smem_optin)
      return fail(CNF_ERR_UNSUPPORTED, "H=%d needs %zu bytes of shared memory (> %d)", dims->H, smem, di.max_smem_optin);
    const int64_t tiles = T * ((P + cnf::kSimtTM - 1) / cnf::kSimtTM);
    if (int rc = cnf::host::ensure_smem(cnf::simt_backward_kernel, smem, di.device, smem_set)) return rc;
    cnf::simt_backward_kernel<<<(unsigned)simt_grid(tiles, di.sms), 256, smem, st>>>(
        *dims, packed, d_gout, static_cast<const float*>(d_stash), d_gshift, T, P);
    CNF_CUDA(cudaGetLastError());
    return CNF_OK;
  }
  if (!tc_ok(*dims))
    return fail(CNF_ERR_UNSUPPORTED, "tensor-core path unsupported for H=%d nl=%d cin=%d cout=%d; use CNF_PREC_FP32",
                dims->H, dims->nl, dims->cin, dims->cout);
  BwdArgs b{*dims, packed, d_gout, d_stash, d_gshift, T, P, st};
  return use_tc2(*dims) ? cnf::host::tc2_backward(b) : cnf::host::tc_backward(b);
}

int cnf_query_launch(const cnf_dims* dims, int precision, int64_t T, int64_t P, int64_t* values, int n) {
  if (int rc = check_dims(dims)) return rc;
  if (!values || n < 1) return fail(CNF_ERR_INVALID_ARGUMENT, "values is NULL or n < 1");
  if (T < 1 || P < 1) return fail(CNF_ERR_INVALID_ARGUMENT, "T=%lld P=%lld", (long long)T, (long long)P);
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  LaunchInfo li;
  if (precision == CNF_PREC_FP32) {
    const int64_t tiles = T * ((P + cnf::kSimtTM - 1) / cnf::kSimtTM);
    li = LaunchInfo{simt_grid(tiles, di.sms), 256, simt_smem_bytes(*dims), 0, 0, cnf::kSimtTM};
  } else {
    if (!is_tc_precision(precision)) return fail(CNF_ERR_INVALID_ARGUMENT, "unknown precision %d", precision);
    if (!tc_ok(*dims)) return fail(CNF_ERR_UNSUPPORTED, "tensor-core path unsupported for these dims");
    FwdArgs a{};
    a.d = *dims;
    a.T = T;
    a.P = P;
    a.query = &li;
    if (int rc = tc_forward_dispatch(precision, a)) return rc;
  }
  const int64_t v[7] = {di.sms, li.grid, li.threads, (int64_t)li.smem, li.ctas_per_sm, li.tmem_cols, li.tile_points};
  for (int i = 0; i < n && i < 7; ++i) values[i] = v[i];
  return CNF_OK;
}

}  // extern "C"
