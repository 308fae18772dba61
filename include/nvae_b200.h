/*
 * nvae_b200.h -- C ABI of libnvae_b200.so: B200 (sm_100a) kernels for the NVAE-TF hot path.
 *
 * This is the drop-in boundary.  Every entry point is an asynchronous,
 * launch-only wrapper on the caller's CUDA stream: plain device pointers and sizes, POD
 * descriptor structs, no C++/torch/TensorFlow types.  The caller (the TF custom-op
 * `Compute()` bodies in tf_op/, or ctypes in nvae_tf_b200/_lib.py) owns every buffer
 * including workspaces; kernels never allocate, never retain pointers past return and never
 * synchronise the host.  Return value: 0 on success, a positive cudaError_t, or a negative
 * NVAE_E_* code.  There is no CPU fallback: an unsupported shape is an error.
 *
 * All tensors are float32; activations NHWC, conv kernels HWIO ([R,S,Cin,Cout]), depthwise
 * kernels [5,5,C,1], dense kernels [in,out] -- TensorFlow's layouts, so the reference's
 * checkpoints and variables map 1:1.  "Replaces" lines cite the reference call sites
 * (file:line in stevensdavid/nvae-tf) whose TensorFlow ops the entry point stands in for.
 */
#ifndef NVAE_B200_H_
#define NVAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* nvae_stream_t;

#if defined(__GNUC__)
#define NVAE_API __attribute__((visibility("default")))
#else
#define NVAE_API
#endif

enum {
  NVAE_OK = 0,
  NVAE_E_BADSHAPE = -1,
  NVAE_E_UNSUPPORTED = -2,
  NVAE_E_WORKSPACE = -3,
  NVAE_E_NULLPTR = -4,
  NVAE_E_DRIVER = -5
};

enum { NVAE_ACT_NONE = 0, NVAE_ACT_SWISH = 1, NVAE_ACT_ELU = 2 };

/* Arithmetic of the convolution GEMMs. FP32 = CUDA-core FFMA; TF32 = one tcgen05 kind::tf32 MMA per operand
 * pair, operands rounded to TF32 by their producers (10-bit mantissa: ~5e-4 per conv, NOT within the 1e-3
 * whole-model parity bound); TF32X3 = split-operand products a_hi*b_hi + a_hi*b_lo + a_lo*b_hi on the tensor cores,
 * fp32-level accuracy (22 significant bits per operand): the default and the mode every parity claim is made in.
 * The split is 3xTF32 inside the kernel for most convolutions and, for the large stride-1 GEMMs (>= 20 GFLOP per
 * launch; NVAE_F16X3=0 turns it off, NVAE_F16X3_MIN_GFLOP moves the threshold), 3xFP16: both operands are scaled by
 * a power of two taken from their absmax (computed into the workspace just before the launch) so the fp16 high and
 * low parts stay in range, multiplied with kind::f16 at twice the TF32 rate, and the epilogue undoes the scales. */
enum { NVAE_PREC_FP32 = 0, NVAE_PREC_TF32 = 1, NVAE_PREC_TF32X3 = 2 };

/* Library / device identification. Returns 100 for sm_100; build id string is static. */
NVAE_API int nvae_version(void);
/* Kernels launched by this library since it was loaded (host-side counter; launches recorded into a CUDA graph
 * count once, at capture).  bench.py derives `gpu_launches` from it. */
NVAE_API uint64_t nvae_launch_count(void);
NVAE_API const char* nvae_build_info(void);
/* CUDA-graph plumbing for the host side (the step is captured by the caller's framework; these only instantiate and
 * launch it): instantiate a captured cudaGraph_t so that the per-node priorities recorded at capture -- main chain on a
 * high-priority stream, weight-gradient side stream at normal priority -- are honoured by the CTA scheduler
 * (cudaGraphInstantiateFlagUseNodePriority; a plain instantiation runs every node at the launch stream's priority).
 * `graph` = cudaGraph_t, `*exec` = cudaGraphExec_t.  Return 0 or a cudaError_t. */
NVAE_API int nvae_graph_instantiate(void* graph, int use_node_priority, void** exec);
NVAE_API int nvae_graph_launch(void* exec, nvae_stream_t stream);
NVAE_API int nvae_graph_destroy(void* exec);

/* ------------------------------------------------------------------------------------------
 * BatchNormalization(momentum=0.05, epsilon=1e-5)      Replaces: common.py:148,166;
 * encoder.py:91,95,102,104; decoder.py:125,129,131,135,139-145; preprocess.py:87;
 * postprocess.py:71,84,107 (cuDNN FusedBatchNormV3 + Eigen swish/ELU in the reference).
 *
 * stat is a [4][C] float block the later kernels consume: mean, invstd, scale=gamma*invstd,
 * shift=beta-mean*scale.  training!=0: batch statistics over rows (=N*H*W), biased variance;
 * moving_mean/var updated in place (Bessel-corrected variance, retain factor `momentum`).
 * training==0: stat is derived from the moving statistics and x is not read.
 * ------------------------------------------------------------------------------------------ */
NVAE_API size_t nvae_bn_ws_bytes(int64_t rows, int C);
NVAE_API int nvae_bn_stats(const float* x, int64_t rows, int C, const float* gamma, const float* beta,
                  float* moving_mean, float* moving_var, int training, float momentum, float eps,
                  float* stat, void* ws, size_t ws_bytes, nvae_stream_t stream);

/* nvae_bn_stats followed by nvae_bn_act_fwd (below) as ONE call: out = act(BN(x)) [-> nearest x2], stat and the moving
 * statistics updated as by nvae_bn_stats.  Training-mode tensors that fit L2 run as a single thread-block-cluster
 * launch (per-cluster DSMEM reduction of the statistics, apply pass re-reads from L1/L2); anything else is the two
 * entry points back to back.  Replaces: BatchNormalization + activation pairs at common.py:166-167,
 * encoder.py:102-104, decoder.py:139,143 (same semantics as the two calls it fuses). */
NVAE_API int nvae_bn_fwd(const float* x, int64_t rows, int C, const float* gamma, const float* beta, float* moving_mean,
                float* moving_var, int training, float momentum, float eps, float* stat, int act, int up_h, int up_w,
                int round_tf32, float* out, void* ws, size_t ws_bytes, nvae_stream_t stream);

/* out = act(x*scale+shift) (stat==NULL: out = act(x)); optional nearest x2 upsample
 * (tf.image.resize "nearest", common.py:168-172: pass up_h=H, up_w=W of x, else 0,0);
 * round_tf32!=0 rounds the result to TF32 (RN) so single-pass kind::tf32 consumes it without truncation bias. */
NVAE_API int nvae_bn_act_fwd(const float* x, int64_t rows, int C, const float* stat, int act, int up_h, int up_w,
                    int round_tf32, float* out, nvae_stream_t stream);

/* Backward of nvae_bn_act_fwd (+ batch-norm backward when stat!=NULL):
 *   g = dout*act'(u)  (summed over the 2x2 replicas when upsampled)
 *   training: dx = scale*(g - mean(g) - xhat*mean(g*xhat));  inference: dx = scale*g
 *   dgamma = sum g*xhat, dbeta = sum g   (written with '=' unless NULL)
 *   dx (+)= res_scale*dres when dres!=NULL;  accumulate!=0 adds into existing dx. */
NVAE_API int nvae_bn_act_bwd(const float* dout, const float* x, int64_t rows, int C, const float* stat, int act, int up_h,
                    int up_w, int training, const float* dres, float res_scale, int accumulate, float* dx,
                    float* dgamma, float* dbeta, void* ws, size_t ws_bytes, nvae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SqueezeExcitation + residual merge.   Replaces: common.py:129-142 fused with the cell
 * tails encoder.py:107, decoder.py:147 (y = 0.1*x + SE(t)) and preprocess.py:107,
 * postprocess.py:58 (y = skip + 0.1*SE(t)):   y = alpha*xres + beta*t*gate[b,c].
 * An optional per-channel affine (stat, e.g. decoder.py:145 batch_norm4 -> se) is applied to
 * t first: t' = t*scale+shift.  pooled/hidden/gate are saved for backward.
 * ------------------------------------------------------------------------------------------ */
NVAE_API int nvae_se_fwd(const float* t, const float* stat, const float* xres, int B, int HW, int C, int hid,
                const float* w1, const float* b1, const float* w2, const float* b2, float alpha, float beta,
                float* pooled, float* hidden, float* gate, float* y, nvae_stream_t stream);
NVAE_API size_t nvae_se_bwd_ws_bytes(int B, int C, int hid);
/* dt is the gradient w.r.t. t' (post-affine); dxres (+)= alpha*dy. */
NVAE_API int nvae_se_bwd(const float* dy, const float* t, const float* stat, int B, int HW, int C, int hid, const float* w1,
                const float* w2, const float* pooled, const float* hidden, const float* gate, float alpha,
                float beta, float* dt, float* dxres, int dxres_accumulate, float* dw1, float* db1, float* dw2,
                float* db2, void* ws, size_t ws_bytes, nvae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * DepthwiseConv2D((5,5), padding="same") with the BN-apply + swish of its input fused into
 * the load.   Replaces: decoder.py:130,141-142 (TF DepthwiseConv2dGPUKernelNHWC + 2 more ops).
 * ------------------------------------------------------------------------------------------ */
NVAE_API int nvae_dwconv5x5_fwd(const float* x, const float* stat, int act, int N, int H, int W, int C, const float* w,
                       const float* bias, float* y, nvae_stream_t stream);
/* da = gradient w.r.t. the ACTIVATED input (feed to nvae_bn_act_bwd). */
NVAE_API int nvae_dwconv5x5_bwd_data(const float* dy, int N, int H, int W, int C, const float* w, float* da,
                            nvae_stream_t stream);
NVAE_API size_t nvae_dwconv5x5_bwd_filter_ws_bytes(int N, int H, int W, int C);
NVAE_API int nvae_dwconv5x5_bwd_filter(const float* x, const float* stat, int act, const float* dy, int N, int H, int W,
                              int C, float* dw, float* dbias, void* ws, size_t ws_bytes, nvae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Per-group latent math.   Replaces: Sampler.call common.py:76-102, softclamp5 util.py:49-50,
 * Sampler.sample common.py:65-68 (epsilon supplied by the caller), the per-element KL and its
 * row sum models.py:197-201, and (optional) calculate_log_p util.py:39-46 / decoder.py:69-71,
 * 84-102.  enc_p/dec_p are the raw conv outputs [B,HW,2L] (mu | log_sigma halves);
 * dec_p==NULL is the z_idx==0 branch (standard-normal prior).
 * ------------------------------------------------------------------------------------------ */
/* dist (nullable): [4][B,HW,L] = enc_mu, enc_sigma, dec_mu, dec_sigma (DistributionParams, common.py:12-17).
 * log_q/log_p (nullable, both or neither) are ACCUMULATED into (+=), as decoder.py:97-102 sums over groups. */
NVAE_API int nvae_latent_fwd(const float* enc_p, const float* dec_p, const float* eps, int B, int HW, int L, float* z,
                    float* kl, float* log_q, float* log_p, float* dist, nvae_stream_t stream);
/* kl_weight points at ONE device float: d(total loss)/d(kl[b]) for this group (beta*coeff_g/B). */
NVAE_API int nvae_latent_bwd(const float* enc_p, const float* dec_p, const float* eps, const float* dz,
                    const float* kl_weight, int B, int HW, int L, float* d_enc_p, float* d_dec_p,
                    nvae_stream_t stream);

/* KL balancing + loss assembly.   Replaces: models.py:121-126, 204-222.
 * kl_all [G,B]; recon [B] (nullable); hyper = device floats {beta, ...} (see nvae_schedule_step).
 * balancing: 1/0 = models.py:204 `if balancing`, -1 = decide on device as train_step does (beta<1).
 * Outputs: kl_weight[G] = beta*c_g/B (nullable; consumed by nvae_latent_bwd), kl_loss[B] =
 * beta*sum_g c_g*kl[g,b], scalars (nullable) [0]=total loss incl. bn_loss, [1]=mean(recon+kl_loss). */
NVAE_API int nvae_loss_assemble(const float* kl_all, const float* recon, const float* bn_loss, const float* alphas,
                       const float* hyper, int balancing, int G, int B, float* kl_weight, float* kl_loss,
                       float* scalars, nvae_stream_t stream);

/* Bernoulli(logits).log_prob(x) summed over H,W,C.   Replaces: models.py:242-250
 * (tfp.distributions.Bernoulli).  logits [B,H,W,Cl] with Cl==C or Cl==1 (broadcast);
 * crop>0 evaluates [crop:H-crop, crop:W-crop] only (models.py:243-245). */
NVAE_API int nvae_bernoulli_ll_fwd(const float* logits, const float* x, int B, int H, int W, int C, int Cl, int crop,
                          float* recon, nvae_stream_t stream);
NVAE_API int nvae_bernoulli_ll_bwd(const float* logits, const float* x, int B, int H, int W, int C, int Cl, float scale,
                          float* dlogits, nvae_stream_t stream);

/* Importance-weighted NLL bound over K attempts.   Replaces: evaluate.py:111-123 (tf.stack + reduce_logsumexp +
 * reduce_mean).  recon/log_q/log_p are [K,B] (row k = attempt k: calculate_recon_loss(crop_output=True), and the
 * log q / log p sums of model(batch, nll=True));  per_sample[B] (nullable) = logsumexp_k(-recon - log_q + log_p) - log K;
 * nll[0] = -mean_b per_sample. */
NVAE_API int nvae_iwae_nll(const float* recon, const float* log_q, const float* log_p, int K, int B, float* per_sample,
                  float* nll, nvae_stream_t stream);

/* BN-gamma infinity-norm regulariser.   Replaces: models.py:252-267 (88 x (abs,max) launches).
 * gamma k lives at params+offsets[k] with sizes[k] elements (device tables). */
NVAE_API int nvae_bn_loss_fwd(const float* params, const int64_t* offsets, const int32_t* sizes, int n, float sr_lambda,
                     float* loss, nvae_stream_t stream);
/* grads+offsets[k] += sr_lambda*sign(gamma)/n_ties at the arg-max entries (tf.reduce_max gradient). */
NVAE_API int nvae_bn_loss_bwd(const float* params, float* grads, const int64_t* offsets, const int32_t* sizes, int n,
                     float sr_lambda, nvae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * tfa.layers.SpectralNormalization(power_iterations=1), all layers in three launches, plus
 * the packing of the conv weights into the layouts the tensor-core kernels consume.
 * Replaces: every SpectralNormalization( call site (common.py:41,57,152,156; encoder.py:12,
 * 61,92,96; decoder.py:110,126,132; preprocess.py:20,46-63,91; postprocess.py:29,78,96).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t w_off;    /* kernel [rows,cout] (HWIO flattened) in `params`            */
  int64_t u_off;    /* u [cout] in `state`                                         */
  int64_t v_off;    /* scratch v_raw [rows] in `ws` (floats)                       */
  int64_t t_off;    /* scratch partial t [n_chunks,cout] in `ws` (floats)          */
  int64_t rnd_off;  /* HWIO operand copy in `pack` (floats) or -1                   */
  int64_t tr_off;   /* transposed operand copy [cout_pad][taps][cin_pad] or -1      */
  int32_t rows, cout, taps, cin, cin_pad, cout_pad;
  int32_t chunk0;   /* first row-chunk index of this layer in the flat chunk list  */
  int32_t n_chunks;
} NvaeSnLayer;
#define NVAE_SN_ROWS_PER_CHUNK 64
/* power_iter!=0: v=l2n(W u), u'=l2n(v W), sigma=(vW).u', W/=sigma, u=u' (in place);
 * power_iter==0: weights left untouched (inference / SN inactive), only (re)packed.
 * pack_exact==0: the operand copies are rounded to TF32 (NVAE_PREC_TF32); !=0: exact fp32 (NVAE_PREC_TF32X3). */
NVAE_API int nvae_spectral_norm(float* params, float* state, float* pack, const NvaeSnLayer* layers_dev, int n_layers,
                       const int32_t* chunk_layer_dev, int n_chunks_total, int power_iter, int pack_exact,
                       float* sigma_out, float* ws, nvae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Conv2D(padding="same").   Replaces: every layers.Conv2D call listed under K1 of SURVEY 2.1
 * (cuDNN fwd / bwd-data / bwd-filter in the reference) incl. tf.concat decoder.py:115
 * (second source x2) and the residual add of encoder.py:16.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t N, H, W, Cin;   /* input tensor x [N,H,W,Cin]                                 */
  int32_t Cin2;           /* channels of the concatenated second source x2 (0 = none)   */
  int32_t Cout, R, S;     /* kernel HWIO [R,S,Cin+Cin2,Cout]                            */
  int32_t stride;         /* 1 or 2 (both dims)                                          */
  int32_t Ho, Wo;         /* output spatial size = ceil(H/stride)                        */
  int32_t pad_t, pad_l;   /* TF SAME padding before (extra padding goes after)           */
  int32_t precision;      /* NVAE_PREC_*                                                 */
  int32_t y_ld, y_off;    /* y / dy / residual are [N,Ho,Wo,y_ld] and this conv owns channels
                             [y_off, y_off+Cout) (tf.concat of preprocess.py:73); y_ld==0: Cout */
  float pre_scale, pre_shift; /* x' = x*pre_scale+pre_shift applied on load (preprocess.py:39) FP32 path only */
} NvaeConvDesc;

NVAE_API size_t nvae_conv2d_ws_bytes(const NvaeConvDesc* d, int which /*0 fwd,1 dgrad,2 wgrad*/);
/* y = conv(x ++ x2, w) + bias (+ residual).  w: HWIO fp32 master; w_tr: the transposed operand copy
 * [Cout][taps][Cin+Cin2] written by nvae_spectral_norm (tensor-core path; may be NULL for NVAE_PREC_FP32). */
NVAE_API int nvae_conv2d_fwd(const NvaeConvDesc* d, const float* x, const float* x2, const float* w, const float* w_tr,
                    const float* bias, const float* residual, float* y, void* ws, size_t ws_bytes,
                    nvae_stream_t stream);
/* dx (+)= dgrad(dy, w) for the first Cin channels, dx2 for the concatenated Cin2 channels.
 * w_rnd: the HWIO operand copy from nvae_spectral_norm (NVAE_PREC_TF32: TF32-rounded; TF32X3: w itself works). */
NVAE_API int nvae_conv2d_dgrad(const NvaeConvDesc* d, const float* dy, const float* w, const float* w_rnd, float* dx,
                      float* dx2, int accumulate, void* ws, size_t ws_bytes, nvae_stream_t stream);
/* dw = wgrad(x ++ x2, dy) (HWIO, '='), dbias = column sums of dy (NULL to skip). */
NVAE_API int nvae_conv2d_wgrad(const NvaeConvDesc* d, const float* x, const float* x2, const float* dy, float* dw,
                      float* dbias, void* ws, size_t ws_bytes, nvae_stream_t stream);

/* The same convolution reading act(BN(x)) instead of x: the BatchNorm apply (x*scale[c]+shift[c], rows 2 and 3 of the
 * [4][Cin] stat block nvae_bn_stats / nvae_bn_fwd wrote) and the activation run inside the tensor-core kernel's operand
 * path (the warps that split the staged tile), forward and backward-filter, so the activated tensor is never written or
 * re-read.  Replaces the BN -> (Swish) -> 1x1 Conv2D pairs: decoder.py:125-127 (batch_norm1 -> conv1), decoder.py:143-144
 * (batch_norm3 + swish -> conv2), postprocess.py:71-73 (bn0 -> cbs1.conv), postprocess.py:84-96 (cbs2 BN + swish -> conv3).
 * Backward-data is nvae_conv2d_dgrad as before (its result is the gradient of the ACTIVATED tensor; nvae_bn_act_bwd takes it
 * from there).  Shapes: 1x1, stride 1, one source, Cin % 32 == 0, NVAE_PREC_TF32X3 below the 3xFP16 threshold --
 * nvae_conv2d_bnact_supported says which; the launchers return NVAE_E_UNSUPPORTED otherwise (there is no second path behind
 * them: the caller then applies the BN with nvae_bn_fwd and calls nvae_conv2d_fwd).  Workspace: nvae_conv2d_ws_bytes(d, 0 / 2).
 * Results are bit-identical to nvae_bn_fwd + nvae_conv2d_fwd / _wgrad.  Measured on B200 the fused pair is SLOWER at the cells'
 * sizes (the operand-splitting warps are those GEMMs' critical path), so the host mirror uses it only under
 * NVAE_FUSE_BN_CONV=1|2. */
NVAE_API int nvae_conv2d_bnact_supported(const NvaeConvDesc* d);
NVAE_API int nvae_conv2d_fwd_bnact(const NvaeConvDesc* d, const float* x, const float* stat, int act, const float* w_tr,
                          const float* bias, const float* residual, float* y, void* ws, size_t ws_bytes,
                          nvae_stream_t stream);
NVAE_API int nvae_conv2d_wgrad_bnact(const NvaeConvDesc* d, const float* x, const float* stat, int act, const float* dy,
                            float* dw, float* dbias, void* ws, size_t ws_bytes, nvae_stream_t stream);

/* Rounds a tensor in place to TF32 (round-to-nearest, ties away: cvt.rna.tf32.f32) so that a tensor-core
 * convolution consumes it without the truncation bias of feeding raw fp32 bits to kind::tf32.  Used on
 * conv operands whose producer did not already round them (gradients arriving at nvae_conv2d_dgrad/wgrad). */
NVAE_API int nvae_round_tf32(float* p, int64_t n, nvae_stream_t stream);
/* 1 when nvae_conv2d_{fwd,dgrad,wgrad} (which = 0,1,2) would run this descriptor on the tcgen05 path. */
NVAE_API int nvae_conv2d_uses_tensor_cores(const NvaeConvDesc* d, int which);
/* Launch plan of the tcgen05 path for this descriptor (pure host function, no device needed; returns 0 and zeros when the
 * descriptor does not take that path).  out[16] = {1, BN (tile columns), M tiles, N tiles, k-units per tile, CTAs,
 * raw ring stages, TMEM A slots, accumulators, 3xFP16 (0/1), accumulators per tile (nsub), two M tiles per CTA (dual),
 * split-K fix-up launch (0/1), dynamic shared memory bytes, workspace bytes >> 10, batch chunks of backward-filter}. */
NVAE_API int nvae_conv2d_plan_info(const NvaeConvDesc* d, int which, int32_t* out);

/* ------------------------------------------------------------------------------------------
 * Optimizer + schedules.   Replaces: optimizers.Adamax + CosineDecay train.py:128-131,
 * models.py:121-122,128-129.  `counters` (device int64[2]) = {warm-up metric (model.steps or .epoch),
 * optimizer iterations}; hyper (device float[8]) = {beta, lr_t = lr/(1-b1^t), lr, t, metric}, computed
 * from the counters BEFORE they advance (advance bit0: ++metric, bit1: ++iterations).  One launch
 * per step, graph-capturable.
 * ------------------------------------------------------------------------------------------ */
NVAE_API int nvae_schedule_step(int64_t* counters, float* hyper, float warmup_iters /*0.3*n_total*/, float lr0,
                       float decay_steps, float b1, int advance, nvae_stream_t stream);
NVAE_API int nvae_adamax(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, float b1, float b2,
                float eps, float grad_scale, nvae_stream_t stream);

/* Small utilities used by the host mirror (all launch-only). */
NVAE_API int nvae_fill(float* p, int64_t n, float value, nvae_stream_t stream);
NVAE_API int nvae_axpby(const float* x, float a, float* y, float b, int64_t n, nvae_stream_t stream); /* y=a*x+b*y */
NVAE_API int nvae_broadcast_rows(const float* src, int64_t row_elems, int B, float* dst, nvae_stream_t stream);
NVAE_API int nvae_reduce_rows(const float* src, int64_t row_elems, int B, float* dst, nvae_stream_t stream);
/* z = mu + eps*(sigma*sigma_scale): Sampler.sample on materialised parameters (common.py:65-68;
 * models.py:140-145 temperature, :175-176 extra draws). */
NVAE_API int nvae_reparam(const float* mu, const float* sigma, const float* eps, float sigma_scale, float* z, int64_t n,
                 nvae_stream_t stream);
/* Philox4x32-10 standard normals (production epsilon for common.py:67; parity runs inject eps). */
NVAE_API int nvae_philox_normal(float* out, int64_t n, uint64_t seed, const int64_t* counters, uint64_t stream_id,
                       nvae_stream_t stream);
/* Bernoulli(logits) images: mode 0 = probs_parameter()/mean() = sigmoid(l); mode 1 = sample().
 * Replaces: models.py:168-174, 185-188. */
NVAE_API int nvae_bernoulli_image(const float* logits, int64_t n, int mode, uint64_t seed, uint64_t stream_id, float* out,
                         nvae_stream_t stream);
NVAE_API int nvae_l2_flush(float* scratch, int64_t n, nvae_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NVAE_B200_H_ */
