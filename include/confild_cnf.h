/*
 * confild_cnf.h -- C ABI of the B200-native CNF decode path (libconfild_cnf.so).
 *
 * Drop-in boundary for ONE hot path of CoNFiLD: decoding the FiLM-modulated SIREN
 * auto-decoder, forward and backward-to-latent.  Every entry point replaces a piece
 * of the reference's PyTorch module (paths relative to the reference checkout):
 *
 *   SIRENAutodecoder_film.forward         ConditionalNeuralField/cnf/nf_networks.py:480-495
 *   BatchLinear.forward                   ConditionalNeuralField/cnf/components.py:64-76
 *   Sine.forward                          ConditionalNeuralField/cnf/components.py:19-25
 *   autograd of the above wrt latents     ConditionalDiffusionGeneration/src/guided_diffusion/condition_methods.py:28-33
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross this boundary.
 *   - every pointer named d_* is a DEVICE pointer owned by the caller (PyTorch owns all
 *     buffers); the library never allocates or frees caller-visible memory.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call
 *     synchronises the device.  The library keeps no mutable global state except the
 *     thread-local last-error string.
 *   - return value: 0 = ok, otherwise a CNF_ERR_* code; cnf_last_error() gives the text.
 *   - all tensors are fp32, contiguous, row-major unless stated otherwise.
 *
 * Notation: cin = coordinate features, L = latent features, H = hidden width,
 * nl = number of hidden (H x H) layers, cout = output features, T = frames (latents),
 * P = query points.  w0 is the sine frequency (reference: initialization.py:5).
 */
#ifndef CONFILD_CNF_H_
#define CONFILD_CNF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNF_ABI_VERSION 3

/* error codes */
#define CNF_OK 0
#define CNF_ERR_INVALID_ARGUMENT 1 /* null pointer, non-positive size, bad enum            */
#define CNF_ERR_UNSUPPORTED 2      /* shape/precision combination has no kernel            */
#define CNF_ERR_CUDA 3             /* a CUDA runtime call or launch failed                 */
#define CNF_ERR_BUFFER_TOO_SMALL 4 /* caller-provided buffer smaller than cnf_*_bytes said */

/* operand precision of the hidden-layer GEMM chain (accumulation is always fp32) */
#define CNF_PREC_FP32 0   /* CUDA-core fp32 FMA chain: exact-order reference on the GPU, any H          */
#define CNF_PREC_BF16X3 1 /* tcgen05, bf16 hi/lo split, 3 MMAs per product (a_hi*w_hi + a_lo*w_hi + a_hi*w_lo) */
#define CNF_PREC_FP16 2   /* tcgen05, single fp16 MMA per product (fast mode; error ~4e-4..1e-3)          */
#define CNF_PREC_F16F8 3  /* tcgen05, fp16 product + two fp8 (kind::f8f6f4) correction products at twice the fp16 rate:
                           *   a*w ~= f16(a)*f16(Sw) + e5m2(a - f16 a)*e4m3(Sw) + e5m2(a)*e4m3(Sw - f16(Sw)),  S = 2^n per layer
                           * = 2 MMA-equivalents per product instead of 3; forward error 2e-5..9e-5.
                           * Forward only: cnf_backward with this precision runs the bf16 hi/lo kernels on the same stash. */

typedef struct cnf_dims {
  int32_t cin;  /* in_coord_features   */
  int32_t L;    /* in_latent_features  */
  int32_t H;    /* hidden_features     */
  int32_t nl;   /* num_hidden_layers   */
  int32_t cout; /* out_features        */
} cnf_dims;

/* Library / ABI version and the text of the last error raised on this thread. */
int cnf_abi_version(void);
const char* cnf_last_error(void);

/* Debug / tuning knobs that select another schedule of the same arithmetic.  Their initial values come from the
 * environment variables of the same name, read once per process; this call overrides one at run time (tests).
 *   "CNF_TC2"        0: route H = 128 through the generic kernels (default 1)
 *   "CNF_TC_STAGES"  n: cap the depth of the shared-memory weight ring (default 0 = as deep as fits)
 *   "CNF_TC_PACKED"  0/1: force frame-aligned / packed tiles (default -1 = by shape)
 *   "CNF_TC_CLUSTER" H = 256/384 forward: 1 = CTA pairs with multicast weight stages (default), 2 = CTA pairs driven by
 *                    cta_group::2 MMAs (correct, measured ~10 % slower: DESIGN.md section 4), 0 = single CTAs
 *   "CNF_GN_CLUSTER" n: cnf_group_norm_nhwc_bf16 runs activations of up to n KiB per sample as ONE launch of an 8-CTA
 *                    cluster per sample (default 256), larger ones as two kernels (statistics, then apply); 0 = always two */
int cnf_set_debug_knob(const char* name, int value);

/* 1 if the tensor-core (tcgen05) kernels exist for these dims, else 0 (CUDA-core fp32 only). */
int cnf_tc_supported(const cnf_dims* dims);

/* Number of fp32 elements of the flat parameter vector expected by cnf_pack_weights:
 * the reference module's state_dict order (nf_networks.py:465-468):
 *   net1.0.weight (H,cin), net1.0.bias (H), net1.i.weight (H,H), net1.i.bias (H) for i=1..nl,
 *   net1.{nl+1}.weight (cout,H), net1.{nl+1}.bias (cout), then net2.i.weight (H,L) for i=0..nl. */
int cnf_param_count(const cnf_dims* dims, size_t* count);

/* Size in bytes of the packed-weight buffer (device) for these dims. */
int cnf_packed_bytes(const cnf_dims* dims, size_t* bytes);

/* Pack the module parameters into the device layout the kernels read: w0 folded into the
 * weights and biases, fp32 copies (plain + transposed) for the CUDA-core path and the
 * FiLM-shift GEMM, and pre-swizzled bf16 hi/lo and fp16 shared-memory images of every
 * hidden layer (forward and transposed/backward) for the tcgen05 path.
 * Replaces: the parameter reads inside BatchLinear.forward (components.py:64-71). */
int cnf_pack_weights(const cnf_dims* dims, const float* d_params_flat, float w0,
                     void* d_packed, size_t packed_bytes, void* stream);

/* FiLM shift for every layer in one GEMM:
 *   d_shift[t, l*H + n] = w0 * ( net1[l].bias[n] + sum_k net2[l].weight[n,k] * latents[t,k] )
 * d_latents (T,L), d_shift (T,(nl+1)*H).  Replaces net2[i](latents) at nf_networks.py:492
 * and the bias add at components.py:74, for all i at once. */
int cnf_film_shift(const cnf_dims* dims, const void* d_packed, const float* d_latents, int64_t T,
                   float* d_shift, void* stream);

/* Bytes of the optional cosine stash written by cnf_forward for a later cnf_backward. */
int cnf_stash_bytes(const cnf_dims* dims, int precision, int64_t T, int64_t P, size_t* bytes);

/* Decode.  out[t,p,:] = net1[nl+1]( h_nl ),  h_l = sin( w0*W_l h_{l-1} + shift[t,l,:] ),
 * h_{-1} = coords[p,:].
 *   d_coords: (P,cin) shared by all frames when coord_frame_stride == 0, else frame t reads
 *             d_coords + t*coord_frame_stride (elements), e.g. P*cin for a (T,P,cin) tensor.
 *   d_shift : output of cnf_film_shift for the same T.
 *   d_out   : (T,P,cout).
 *   d_stash : NULL for inference; otherwise cnf_stash_bytes() bytes receiving cos(.) of every
 *             sine argument (fp16 for the tcgen05 precisions, fp32 for CNF_PREC_FP32).
 * Replaces the layer loop nf_networks.py:491-494. */
int cnf_forward(const cnf_dims* dims, const void* d_packed, int precision, const float* d_coords,
                int64_t coord_frame_stride, const float* d_shift, float* d_out, int64_t T, int64_t P,
                void* d_stash, size_t stash_bytes, void* stream);

/* Decode with the all-gather fused into the kernel: this rank's (T,P,cout) block is stored into n_out buffers, the
 * pointers in the HOST array d_outs (1 <= n_out <= 8): its own gathered buffer and the peers' buffers mapped into this
 * process over NVLink (CUDA IPC / torch symmetric memory), each already offset to this rank's frame range.  The epilogue
 * writes 4*cout bytes per point to every target, so the gather costs no extra pass over the field; the caller
 * synchronises the ranks afterwards.  Tensor-core precisions only, no stash.
 * Replaces: decode followed by the one collective of the path, all_gather_into_tensor of the decoded field
 * (the reference itself decodes on a single GPU, inference_function.py:51-76). */
int cnf_forward_gather(const cnf_dims* dims, const void* d_packed, int precision, const float* d_coords,
                       int64_t coord_frame_stride, const float* d_shift, float* const* d_outs, int n_out, int64_t T,
                       int64_t P, void* stream);

/* Backward to the FiLM shifts: given d_gout = dLoss/dout (T,P,cout) and the stash of the
 * matching cnf_forward call, accumulates d_gshift[t, l*H+n] = sum_p dLoss/d(arg of sine l,n at t,p).
 * d_gshift (T,(nl+1)*H) is zeroed by this call before accumulation.
 * Replaces autograd through nf_networks.py:491-494 (condition_methods.py:32). */
int cnf_backward(const cnf_dims* dims, const void* d_packed, int precision, const float* d_gout,
                 const void* d_stash, size_t stash_bytes, float* d_gshift, int64_t T, int64_t P,
                 void* stream);

/* Backward of cnf_film_shift: d_glatents[t,k] = sum_{l,n} d_gshift[t,l*H+n] * w0 * net2[l].weight[n,k].
 * d_glatents (T,L). */
int cnf_film_shift_backward(const cnf_dims* dims, const void* d_packed, const float* d_gshift, int64_t T,
                            float* d_glatents, void* stream);

/* Same, with every output multiplied by the DEVICE scalar *d_scale (NULL = 1): the 1/||r|| of the fused
 * measurement loss below, so that no host round trip or elementwise pass separates the kernels of a DPS step. */
int cnf_film_shift_backward_scaled(const cnf_dims* dims, const void* d_packed, const float* d_gshift, int64_t T,
                                   const float* d_scale, float* d_glatents, void* stream);

/* Fused DPS measurement distance.  Replaces, for one guided-sampling step, the elementwise PyTorch passes between
 * decode and autograd in the reference:
 *     difference = measurement - operator.forward(x0_hat)        guided_diffusion/condition_methods.py:30
 *                  (operator.forward = y_normalizer.denormalize(decode) [* mask])   measurements.py:91-97,222-226
 *     norm = torch.linalg.norm(difference)                        condition_methods.py:31
 *     autograd.grad(norm, x_prev)  -> dnorm/dy = -difference/norm condition_methods.py:32
 * With y_phys = y_scale*y + y_offset (the output normaliser as an affine map per channel) and
 * r = y_meas - mask*y_phys (the operator returns mask*phy_fields; pass an already masked measurement to obtain
 * mask*(y_meas - y_phys)), cnf_forward_loss decodes like cnf_forward and, in the same kernel's head, writes
 *     d_gy[t,p,o]  = -y_scale[o]*mask*r[t,p,o]        (= ||r|| * dnorm/dy, the seed of cnf_backward)
 *     d_norm[0]    = ||r||_2 over all (t,p,o),  d_norm[1] = 1/||r||_2 (0 if ||r|| == 0)
 * so a DPS step is cnf_film_shift -> cnf_forward_loss -> cnf_backward(d_gy) -> cnf_film_shift_backward_scaled(d_norm+1).
 * d_out may be NULL for the tensor-core precisions when the decoded field itself is not needed. */
#define CNF_LOSS_PARTIALS 4096
typedef struct cnf_sensor_loss {
  const float* d_y_meas; /* (T,P,cout) measurement, physical units                                        */
  const float* d_mask;   /* NULL = no mask; else weights of kind mask_kind                               */
  int32_t mask_kind;     /* 1: (P) per point, shared by frames and channels; 2: (P,cout); 3: (T,P,cout)  */
  float y_scale[4];      /* y_phys = y_scale[o]*y + y_offset[o]; identity = 1, 0                          */
  float y_offset[4];
  float* d_gy;           /* (T,P,cout) out                                                                */
  float* d_partials;     /* CNF_LOSS_PARTIALS floats of scratch (per-warp partial sums of r^2)            */
  float* d_norm;         /* 2 floats out                                                                  */
  const float* d_extra_sq; /* NULL, or one device float added to sum r^2 before the square root: the energy
                            * sum y_meas^2 of rows the caller did not pass because their mask weight is zero
                            * (r = y_meas there whatever the decoder returns)                             */
} cnf_sensor_loss;
int cnf_forward_loss(const cnf_dims* dims, const void* d_packed, int precision, const float* d_coords,
                     int64_t coord_frame_stride, const float* d_shift, float* d_out, int64_t T, int64_t P,
                     void* d_stash, size_t stash_bytes, const cnf_sensor_loss* loss, void* stream);

/* Introspection for the bench / tests: fills up to `n` int64 values:
 *   [0] SM count of the current device, [1] CTAs launched by cnf_forward for (precision,T,P),
 *   [2] threads per CTA, [3] dynamic shared memory bytes per CTA, [4] resident CTAs per SM,
 *   [5] TMEM columns per CTA (0 for the CUDA-core path), [6] points per tile. */
int cnf_query_launch(const cnf_dims* dims, int precision, int64_t T, int64_t P, int64_t* values, int n);

/* ---- sampler-side helper (not part of the decode path) ------------------------------------------
 * GroupNorm over channels-last bf16 activations with fp32 statistics, an optional per-(sample, channel) value added
 * before the normalisation (the residual block's timestep-embedding add) and an optional SiLU:
 *     y[n,p,c] = act( ((x[n,p,c] + add[n,c]) - mean[n,g]) * rstd[n,g] * gamma[c] + beta[c] ),   g = c / (C/groups)
 * x, y: (N, HW, C) bf16 = a torch channels_last (N,C,H,W) tensor, 16-byte aligned, C a multiple of 8 and of `groups`
 * (<= 64); d_add NULL or fp32 rows of C values, row n at d_add + n * add_stride (a slice of a wider matrix);
 * d_partials: scratch of cnf_group_norm_scratch_bytes(N) bytes.  Deterministic
 * (no atomics).  Replaces GroupNorm32 + SiLU of the guided-diffusion U-Net's blocks
 * (UnconditionalDiffusionTraining_and_Generation/src/nn.py:17-19, src/unet.py:185-200,228-256,283-300) in the
 * inference-only fast path of confild_b200.LatentUNet. */
#define CNF_GN_MAX_CHUNKS 128
size_t cnf_group_norm_scratch_bytes(int64_t N);
int cnf_group_norm_nhwc_bf16(const void* d_x, const float* d_add, int64_t add_stride, const float* d_gamma,
                             const float* d_beta, void* d_y, float* d_partials, int64_t N, int64_t HW, int32_t C,
                             int32_t groups, float eps, int32_t silu, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CONFILD_CNF_H_ */
