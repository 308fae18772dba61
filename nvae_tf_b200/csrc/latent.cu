// Per-group latent math, KL balancing / loss assembly, Bernoulli reconstruction likelihood and the
// BN-gamma regulariser.  Reference: common.py:65-102, util.py:39-50, models.py:191-267.
// All of these move <= a few MB per launch (launch-latency bound at the model's sizes): one CTA per
// sample, deterministic fixed-tree reductions, everything the reference does in ~25 TF ops per
// group fused into one pass.
#include "common.cuh"

namespace nvae {

constexpr int kLatThreads = 256;
constexpr float kHalfLog2Pi = 0.91893853320467274178f;

__device__ __forceinline__ float sc5(float x) { return 5.f * tanhf(x * 0.2f); }
__device__ __forceinline__ float sc5_grad_from_val(float y) {  // d/dx 5*tanh(x/5) = 1 - tanh^2 = 1 - (y/5)^2
  const float t = y * 0.2f;
  return 1.f - t * t;
}

struct LatentVals {
  float mu_q, sig_q, mu_p, sig_p;
  float gA, gB, gC, gD;  // softclamp derivatives at (a+c), (b+d), c, d
};

__device__ __forceinline__ LatentVals latent_eval(float a, float b, bool has_dec, float c, float d) {
  LatentVals v;
  if (has_dec) {
    const float yA = sc5(a + c), yB = sc5(b + d), yC = sc5(c), yD = sc5(d);
    v.mu_q = yA; v.sig_q = expf(yB) + 1e-2f; v.mu_p = yC; v.sig_p = expf(yD) + 1e-2f;
    v.gA = sc5_grad_from_val(yA); v.gB = sc5_grad_from_val(yB); v.gC = sc5_grad_from_val(yC);
    v.gD = sc5_grad_from_val(yD);
  } else {
    const float yA = sc5(a), yB = sc5(b);
    v.mu_q = yA; v.sig_q = expf(yB) + 1e-2f; v.mu_p = 0.f; v.sig_p = 1.f;
    v.gA = sc5_grad_from_val(yA); v.gB = sc5_grad_from_val(yB); v.gC = 0.f; v.gD = 0.f;
  }
  return v;
}

__global__ void __launch_bounds__(kLatThreads) latent_fwd_kernel(const float* __restrict__ enc_p,
                                                                 const float* __restrict__ dec_p,
                                                                 const float* __restrict__ eps, int HW, int L,
                                                                 float* __restrict__ z, float* __restrict__ kl,
                                                                 float* __restrict__ log_q, float* __restrict__ log_p,
                                                                 float* __restrict__ dist, int64_t dist_stride) {
  nvae::pdl_enter();
  __shared__ float red[33];
  const int b = blockIdx.x, n = HW * L;
  const int64_t base = (int64_t)b * n;
  float s_kl = 0.f, s_q = 0.f, s_p = 0.f;
  for (int i = threadIdx.x; i < n; i += kLatThreads) {
    const int p = i / L, l = i - p * L;
    const int64_t o = ((int64_t)b * HW + p) * 2 * L + l;
    const float a = enc_p[o], bb = enc_p[o + L];
    float c = 0.f, d = 0.f;
    if (dec_p != nullptr) { c = dec_p[o]; d = dec_p[o + L]; }
    const LatentVals v = latent_eval(a, bb, dec_p != nullptr, c, d);
    const float e = eps[base + i];
    const float zz = fmaf(e, v.sig_q, v.mu_q);
    z[base + i] = zz;
    if (dist != nullptr) {  // DistributionParams of common.py:12-17
      dist[base + i] = v.mu_q; dist[dist_stride + base + i] = v.sig_q;
      dist[2 * dist_stride + base + i] = v.mu_p; dist[3 * dist_stride + base + i] = v.sig_p;
    }
    const float t1 = (v.mu_q - v.mu_p) / v.sig_p, t2 = v.sig_q / v.sig_p;
    s_kl += 0.5f * (t1 * t1 + t2 * t2) - 0.5f - logf(t2);
    if (log_q != nullptr) {
      s_q += -0.5f * e * e - kHalfLog2Pi - logf(v.sig_q);  // (z-mu_q)/sig_q == eps
      const float nz = (zz - v.mu_p) / v.sig_p;
      s_p += -0.5f * nz * nz - kHalfLog2Pi - logf(v.sig_p);
    }
  }
  s_kl = block_sum(s_kl, red);
  if (threadIdx.x == 0) kl[b] = s_kl;
  if (log_q != nullptr) {
    s_q = block_sum(s_q, red);
    s_p = block_sum(s_p, red);
    if (threadIdx.x == 0) {
      log_q[b] += s_q;  // accumulated over groups (decoder.py:97-102); caller zeroes them once
      log_p[b] += s_p;
    }
  }
}

__global__ void latent_bwd_kernel(const float* __restrict__ enc_p, const float* __restrict__ dec_p,
                                  const float* __restrict__ eps, const float* __restrict__ dz,
                                  const float* __restrict__ kl_weight, int64_t total, int L,
                                  float* __restrict__ d_enc_p, float* __restrict__ d_dec_p) {
  nvae::pdl_enter();
  const float w = kl_weight[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / L;
    const int l = (int)(i - p * L);
    const int64_t o = p * 2 * L + l;
    const float a = enc_p[o], bb = enc_p[o + L];
    float c = 0.f, d = 0.f;
    if (dec_p != nullptr) { c = dec_p[o]; d = dec_p[o + L]; }
    const LatentVals v = latent_eval(a, bb, dec_p != nullptr, c, d);
    const float e = eps[i], g = dz != nullptr ? dz[i] : 0.f;
    const float inv_sp = 1.f / v.sig_p;
    const float t1 = (v.mu_q - v.mu_p) * inv_sp, t2 = v.sig_q * inv_sp;
    const float d_muq = g + w * t1 * inv_sp;
    const float d_sigq = g * e + w * (t2 * inv_sp - 1.f / v.sig_q);
    const float dA = d_muq * v.gA;
    const float dB = d_sigq * (v.sig_q - 1e-2f) * v.gB;
    d_enc_p[o] = dA;
    d_enc_p[o + L] = dB;
    if (dec_p != nullptr) {
      const float d_mup = -w * t1 * inv_sp;
      const float d_sigp = w * (1.f - t1 * t1 - t2 * t2) * inv_sp;
      d_dec_p[o] = dA + d_mup * v.gC;
      d_dec_p[o + L] = dB + d_sigp * (v.sig_p - 1e-2f) * v.gD;
    }
  }
}

// Single CTA.  models.py:204-222 (balancing when beta<1) and models.py:121-126.
__global__ void __launch_bounds__(256) loss_assemble_kernel(const float* __restrict__ kl_all,
                                                            const float* __restrict__ recon,
                                                            const float* __restrict__ bn_loss,
                                                            const float* __restrict__ alphas,
                                                            const float* __restrict__ hyper, int balancing_mode,
                                                            int G, int B, float* __restrict__ kl_weight,
                                                            float* __restrict__ kl_loss,
                                                            float* __restrict__ scalars) {
  nvae::pdl_enter();
  __shared__ float red[33];
  __shared__ float coeff[1024];
  const float beta = hyper[0];
  const bool balancing = balancing_mode < 0 ? beta < 1.f : balancing_mode != 0;  // models.py:123
  float total = 0.f;
  for (int g = 0; g < G; ++g) {
    float s = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) s += fabsf(kl_all[(int64_t)g * B + b]);
    s = block_sum(s, red);
    const float c = s / (float)B + 0.01f;
    if (threadIdx.x == 0) coeff[g] = c;
    total += c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float mean = 0.f;
    for (int g = 0; g < G; ++g) {
      coeff[g] = balancing ? coeff[g] / alphas[g] * total : 1.f;
      mean += coeff[g];
    }
    mean /= (float)G;
    for (int g = 0; g < G; ++g) {
      if (balancing) coeff[g] /= mean;
      if (kl_weight != nullptr) kl_weight[g] = beta * coeff[g] / (float)B;
    }
  }
  __syncthreads();
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float s = 0.f;
    for (int g = 0; g < G; ++g) s = fmaf(coeff[g], kl_all[(int64_t)g * B + b], s);
    s *= beta;
    kl_loss[b] = s;
    acc += s + (recon != nullptr ? recon[b] : 0.f);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0 && scalars != nullptr) {
    const float m = acc / (float)B;
    scalars[1] = m;
    scalars[0] = m + (bn_loss ? bn_loss[0] : 0.f);
  }
}

__device__ __forceinline__ float softplusf(float l) { return fmaxf(l, 0.f) + log1pf(expf(-fabsf(l))); }

__global__ void __launch_bounds__(256) bernoulli_fwd_kernel(const float* __restrict__ logits,
                                                            const float* __restrict__ x, int H, int W, int C, int Cl,
                                                            int crop, float* __restrict__ recon) {
  nvae::pdl_enter();
  __shared__ float red[33];
  const int b = blockIdx.x;
  const int Hc = H - 2 * crop, Wc = W - 2 * crop, n = Hc * Wc * C;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i % C, t = i / C, w = t % Wc + crop, h = t / Wc + crop;
    const int64_t pix = ((int64_t)b * H + h) * W + w;
    const float l = logits[pix * Cl + (Cl == 1 ? 0 : c)];
    const float xv = x[pix * C + c];
    s += xv * l - softplusf(l);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) recon[b] = -s;
}

// dlogits = scale * (sigmoid(l) - x), summed over the broadcast channel when Cl==1
__global__ void bernoulli_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ x, int64_t npix, int C,
                                     int Cl, float scale, float* __restrict__ dlogits) {
  nvae::pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix * Cl; i += (int64_t)gridDim.x * blockDim.x) {
    const float l = logits[i];
    const float sg = 1.f / (1.f + expf(-l));
    float g;
    if (Cl == C) {
      g = sg - x[i];
    } else {
      g = 0.f;
      for (int c = 0; c < C; ++c) g += sg - x[i * C + c];
    }
    dlogits[i] = scale * g;
  }
}

// One CTA of 32 warps; warp w handles layers w, w+32, ...; fixed-order final sum.
__global__ void __launch_bounds__(1024) bn_loss_kernel(const float* __restrict__ params, float* __restrict__ grads,
                                                       const int64_t* __restrict__ offsets,
                                                       const int32_t* __restrict__ sizes, int n, float lambda,
                                                       float* __restrict__ loss) {
  nvae::pdl_enter();
  __shared__ float smax[1024];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = warp; k < n; k += 32) {
    const float* g = params + offsets[k];
    const int sz = sizes[k];
    float m = 0.f;
    for (int i = lane; i < sz; i += 32) m = fmaxf(m, fabsf(g[i]));
    m = warp_max(m);
    if (lane == 0) smax[k] = m;
    if (grads != nullptr) {  // tf.reduce_max gradient: split evenly among ties
      int ties = 0;
      for (int i = lane; i < sz; i += 32) ties += (fabsf(g[i]) == m);
      for (int o = 16; o > 0; o >>= 1) ties += __shfl_xor_sync(0xffffffffu, ties, o);
      float* dg = grads + offsets[k];
      const float q = lambda / (float)ties;
      for (int i = lane; i < sz; i += 32)
        if (fabsf(g[i]) == m) dg[i] += g[i] > 0.f ? q : (g[i] < 0.f ? -q : 0.f);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && loss != nullptr) {
    float s = 0.f;
    for (int k = 0; k < n; ++k) s += smax[k];
    loss[0] = lambda * s;
  }
}

}  // namespace nvae

using namespace nvae;

extern "C" int nvae_latent_fwd(const float* enc_p, const float* dec_p, const float* eps, int B, int HW, int L,
                               float* z, float* kl, float* log_q, float* log_p, float* dist, nvae_stream_t stream) {
  if (B <= 0 || HW <= 0 || L <= 0) return NVAE_E_BADSHAPE;
  if (!enc_p || !eps || !z || !kl) return NVAE_E_NULLPTR;
  if ((log_q == nullptr) != (log_p == nullptr)) return NVAE_E_NULLPTR;
  nvae::launch(latent_fwd_kernel, B, kLatThreads, 0, stream, enc_p, dec_p, eps, HW, L, z, kl, log_q, log_p, dist,
                                                   (int64_t)B * HW * L);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_latent_bwd(const float* enc_p, const float* dec_p, const float* eps, const float* dz,
                               const float* kl_weight, int B, int HW, int L, float* d_enc_p, float* d_dec_p,
                               nvae_stream_t stream) {
  if (B <= 0 || HW <= 0 || L <= 0) return NVAE_E_BADSHAPE;
  if (!enc_p || !eps || !kl_weight || !d_enc_p) return NVAE_E_NULLPTR;
  if (dec_p != nullptr && d_dec_p == nullptr) return NVAE_E_NULLPTR;
  const int64_t total = (int64_t)B * HW * L;
  int64_t grid = ceil_div(total, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  nvae::launch(latent_bwd_kernel, (int)grid, 256, 0, stream, enc_p, dec_p, eps, dz, kl_weight, total, L, d_enc_p, d_dec_p);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_loss_assemble(const float* kl_all, const float* recon, const float* bn_loss, const float* alphas,
                                  const float* hyper, int balancing, int G, int B, float* kl_weight, float* kl_loss,
                                  float* scalars, nvae_stream_t stream) {
  if (G <= 0 || G > 1024 || B <= 0) return NVAE_E_BADSHAPE;
  if (!kl_all || !alphas || !hyper || !kl_loss) return NVAE_E_NULLPTR;
  nvae::launch(loss_assemble_kernel, 1, 256, 0, stream, kl_all, recon, bn_loss, alphas, hyper, balancing, G, B, kl_weight, kl_loss,
                                              scalars);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_bernoulli_ll_fwd(const float* logits, const float* x, int B, int H, int W, int C, int Cl, int crop,
                                     float* recon, nvae_stream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || (Cl != C && Cl != 1) || crop < 0 || 2 * crop >= H || 2 * crop >= W)
    return NVAE_E_BADSHAPE;
  if (!logits || !x || !recon) return NVAE_E_NULLPTR;
  nvae::launch(bernoulli_fwd_kernel, B, 256, 0, stream, logits, x, H, W, C, Cl, crop, recon);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_bernoulli_ll_bwd(const float* logits, const float* x, int B, int H, int W, int C, int Cl,
                                     float scale, float* dlogits, nvae_stream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || (Cl != C && Cl != 1)) return NVAE_E_BADSHAPE;
  if (!logits || !x || !dlogits) return NVAE_E_NULLPTR;
  const int64_t npix = (int64_t)B * H * W;
  int64_t grid = ceil_div(npix * Cl, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  nvae::launch(bernoulli_bwd_kernel, (int)grid, 256, 0, stream, logits, x, npix, C, Cl, scale, dlogits);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

// Importance-weighted bound (evaluate.py:111-123): log_iw[k,b] = -recon[k,b] - log_q[k,b] + log_p[k,b];
// per_sample[b] = logsumexp_k log_iw[k,b] - log K;  nll = -mean_b per_sample.  One block, fixed-order tree: deterministic.
namespace nvae {
__global__ void __launch_bounds__(256) iwae_nll_kernel(const float* __restrict__ recon, const float* __restrict__ log_q,
                                                       const float* __restrict__ log_p, int K, int B,
                                                       float* __restrict__ per_sample, float* __restrict__ nll) {
  nvae::pdl_enter();
  __shared__ float red[256];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) {
      const int64_t i = (int64_t)k * B + b;
      m = fmaxf(m, -recon[i] - log_q[i] + log_p[i]);
    }
    float s = 0.f;
    for (int k = 0; k < K; ++k) {
      const int64_t i = (int64_t)k * B + b;
      s += expf(-recon[i] - log_q[i] + log_p[i] - m);
    }
    const float v = m + logf(s) - logf((float)K);
    if (per_sample) per_sample[b] = v;
    acc += v;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) nll[0] = -red[0] / (float)B;
}
}  // namespace nvae

extern "C" int nvae_iwae_nll(const float* recon, const float* log_q, const float* log_p, int K, int B, float* per_sample,
                             float* nll, nvae_stream_t stream) {
  if (K <= 0 || B <= 0) return NVAE_E_BADSHAPE;
  if (!recon || !log_q || !log_p || !nll) return NVAE_E_NULLPTR;
  nvae::launch(nvae::iwae_nll_kernel, 1, 256, 0, stream, recon, log_q, log_p, K, B, per_sample, nll);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_bn_loss_fwd(const float* params, const int64_t* offsets, const int32_t* sizes, int n,
                                float sr_lambda, float* loss, nvae_stream_t stream) {
  if (n <= 0 || n > 1024) return NVAE_E_BADSHAPE;
  if (!params || !offsets || !sizes || !loss) return NVAE_E_NULLPTR;
  nvae::launch(bn_loss_kernel, 1, 1024, 0, stream, params, nullptr, offsets, sizes, n, sr_lambda, loss);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_bn_loss_bwd(const float* params, float* grads, const int64_t* offsets, const int32_t* sizes, int n,
                                float sr_lambda, nvae_stream_t stream) {
  if (n <= 0 || n > 1024) return NVAE_E_BADSHAPE;
  if (!params || !grads || !offsets || !sizes) return NVAE_E_NULLPTR;
  nvae::launch(bn_loss_kernel, 1, 1024, 0, stream, params, grads, offsets, sizes, n, sr_lambda, nullptr);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}
