// Generic fp32 CUDA-core implicit-GEMM convolution: forward, backward-data, backward-filter.
// This is the NVAE_PREC_FP32 arithmetic mode (exact-fp32 FFMA, used for bit-tight parity runs) and
// the path for the shapes the tcgen05 kernels do not take (stride 2, Cin=1 stem, Cout=1 head).
// Any stride in {1,2}, any channel count, TF SAME padding, concatenated second source, fused
// input affine (2x-1 of preprocess.py:39), bias, residual add.  64x64x16 tiles, 4x4 per thread.
#include "common.cuh"

namespace nvae {

constexpr int kBM = 64, kBN = 64, kBK = 16, kSimtThreads = 256;

struct ConvGeom {
  int N, H, W, Cin, Cin2, Ct, Cout, R, S, stride, Ho, Wo, pad_t, pad_l, y_ld, y_off;
  float pre_scale, pre_shift;
};

__device__ __forceinline__ float conv_load_x(const ConvGeom& g, const float* __restrict__ x,
                                             const float* __restrict__ x2, int n, int h, int w, int c) {
  if (h < 0 || h >= g.H || w < 0 || w >= g.W) return 0.f;
  const int64_t pix = ((int64_t)n * g.H + h) * g.W + w;
  const float v = c < g.Cin ? __ldg(x + pix * g.Cin + c) : __ldg(x2 + pix * g.Cin2 + (c - g.Cin));
  return fmaf(v, g.pre_scale, g.pre_shift);
}

// ---- forward: M=(n,ho,wo) K=(r,s,c) N=co -----------------------------------------------------
struct FwdProb {
  ConvGeom g;
  const float *x, *x2, *w, *bias, *res;
  float* y;
  int64_t M;
  int K, Nn;
  static constexpr bool kAKFast = true, kBKFast = false;
  struct Row { int n, ho, wo; bool ok; };
  __device__ Row row(int64_t m) const {
    Row r;
    r.ok = m < M;
    const int64_t mm = r.ok ? m : 0;
    r.wo = (int)(mm % g.Wo);
    const int64_t t = mm / g.Wo;
    r.ho = (int)(t % g.Ho);
    r.n = (int)(t / g.Ho);
    return r;
  }
  __device__ float loadA(const Row& r, int k) const {
    if (!r.ok || k >= K) return 0.f;
    const int c = k % g.Ct, tap = k / g.Ct, s = tap % g.S, rr = tap / g.S;
    return conv_load_x(g, x, x2, r.n, r.ho * g.stride + rr - g.pad_t, r.wo * g.stride + s - g.pad_l, c);
  }
  __device__ float loadB(int k, int n) const { return (k < K && n < Nn) ? __ldg(w + (int64_t)k * g.Cout + n) : 0.f; }
  __device__ void store(int64_t m, int n, float v) const {
    if (m >= M || n >= Nn) return;
    if (bias) v += __ldg(bias + n);
    if (res) v += __ldg(res + m * g.y_ld + g.y_off + n);
    y[m * g.y_ld + g.y_off + n] = v;
  }
};

// ---- backward data: M=(n,h,w) K=(r,s,co) N=ci ------------------------------------------------
struct DgradProb {
  ConvGeom g;
  const float *dy, *w;
  float *dx, *dx2;
  int accumulate;
  int64_t M;
  int K, Nn;
  static constexpr bool kAKFast = true, kBKFast = true;
  struct Row { int n, h, w; bool ok; };
  __device__ Row row(int64_t m) const {
    Row r;
    r.ok = m < M;
    const int64_t mm = r.ok ? m : 0;
    r.w = (int)(mm % g.W);
    const int64_t t = mm / g.W;
    r.h = (int)(t % g.H);
    r.n = (int)(t / g.H);
    return r;
  }
  __device__ float loadA(const Row& r, int k) const {
    if (!r.ok || k >= K) return 0.f;
    const int co = k % g.Cout, tap = k / g.Cout, s = tap % g.S, rr = tap / g.S;
    const int hh = r.h + g.pad_t - rr, ww = r.w + g.pad_l - s;
    if (hh < 0 || ww < 0) return 0.f;
    if (g.stride == 2 && ((hh | ww) & 1)) return 0.f;
    const int ho = hh / g.stride, wo = ww / g.stride;
    if (ho >= g.Ho || wo >= g.Wo) return 0.f;
    return __ldg(dy + (((int64_t)r.n * g.Ho + ho) * g.Wo + wo) * g.y_ld + g.y_off + co);
  }
  __device__ float loadB(int k, int n) const {
    if (k >= K || n >= Nn) return 0.f;
    const int co = k % g.Cout, tap = k / g.Cout;
    return __ldg(w + ((int64_t)tap * g.Ct + n) * g.Cout + co);
  }
  __device__ void store(int64_t m, int n, float v) const {
    if (m >= M || n >= Nn) return;
    float* p = n < g.Cin ? dx + m * g.Cin + n : dx2 + m * g.Cin2 + (n - g.Cin);
    if (p == nullptr) return;
    *p = accumulate ? *p + v : v;
  }
};

// ---- backward filter: M=(r,s,c) K=(n,ho,wo) N=co, split-K over grid.z --------------------------
struct WgradProb {
  ConvGeom g;
  const float *x, *x2, *dy;
  float* dw;
  int64_t Kpix;  // N*Ho*Wo
  int M, Nn, split;  // split!=0: dw is a [splits][M][Nn] partial buffer reduced in fixed order afterwards
  static constexpr bool kAKFast = false, kBKFast = false;
  struct Row { int c, r, s; bool ok; };
  __device__ Row row(int64_t m) const {
    Row q;
    q.ok = m < M;
    const int mm = q.ok ? (int)m : 0;
    q.c = mm % g.Ct;
    const int tap = mm / g.Ct;
    q.s = tap % g.S;
    q.r = tap / g.S;
    return q;
  }
  __device__ float loadA(const Row& q, int64_t k) const {
    if (!q.ok || k >= Kpix) return 0.f;
    const int wo = (int)(k % g.Wo);
    const int64_t t = k / g.Wo;
    const int ho = (int)(t % g.Ho), n = (int)(t / g.Ho);
    return conv_load_x(g, x, x2, n, ho * g.stride + q.r - g.pad_t, wo * g.stride + q.s - g.pad_l, q.c);
  }
  __device__ float loadB(int64_t k, int n) const {
    return (k < Kpix && n < Nn) ? __ldg(dy + k * g.y_ld + g.y_off + n) : 0.f;
  }
  __device__ void store(int64_t m, int n, float v) const {
    if (m >= M || n >= Nn) return;
    if (split) dw[((int64_t)blockIdx.z * M + m) * Nn + n] = v;
    else dw[m * g.Cout + n] = v;
  }
};

template <class Prob>
__global__ void __launch_bounds__(kSimtThreads) simt_conv_kernel(const Prob p, int64_t k_total, int64_t k_per_split) {
  nvae::pdl_enter();
  __shared__ float As[kBK][kBM + 4];
  __shared__ float Bs[kBK][kBN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * kBM;
  const int n0 = blockIdx.y * kBN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
  const int64_t kend = kbeg + k_per_split < k_total ? kbeg + k_per_split : k_total;
  typename Prob::Row rows[4];
  if (Prob::kAKFast) {
#pragma unroll
    for (int i = 0; i < 4; ++i) rows[i] = p.row(m0 + (tid >> 4) + i * 16);
  } else {
    rows[0] = p.row(m0 + (tid & 63));
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t k0 = kbeg; k0 < kend; k0 += kBK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (Prob::kAKFast) {
        const int kk = tid & 15, mm = (tid >> 4) + i * 16;
        const int64_t k = k0 + kk;
        As[kk][mm] = k < kend ? p.loadA(rows[i], k) : 0.f;
      } else {
        const int mm = tid & 63, kk = (tid >> 6) + i * 4;
        const int64_t k = k0 + kk;
        As[kk][mm] = k < kend ? p.loadA(rows[0], k) : 0.f;
      }
      if (Prob::kBKFast) {
        const int kk = tid & 15, nn = (tid >> 4) + i * 16;
        const int64_t k = k0 + kk;
        Bs[kk][nn] = k < kend ? p.loadB(k, n0 + nn) : 0.f;
      } else {
        const int nn = tid & 63, kk = (tid >> 6) + i * 4;
        const int64_t k = k0 + kk;
        Bs[kk][nn] = k < kend ? p.loadB(k, n0 + nn) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) p.store(m0 + ty * 4 + i, n0 + tx * 4 + j, acc[i][j]);
}

// column sums of a [rows, C] matrix (bias gradient): each CTA sums a row range of 32 columns into part[split][C]; the last
// CTA of a column chunk to finish (ticket[blockIdx.x], zeroed ahead of the launch) adds the chunk's partials in split order --
// deterministic, one launch
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ a, int64_t rows, int C, int ld,
                                                     int64_t rows_per_split, float* __restrict__ part,
                                                     unsigned* __restrict__ ticket, float* __restrict__ out) {
  nvae::pdl_enter();
  __shared__ float sm[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r1 = r0 + rows_per_split < rows ? r0 + rows_per_split : rows;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < C) {
    int64_t r = r0 + ry;
    for (; r + 24 < r1; r += 32) {  // four independent loads in flight, fixed association
      s0 += __ldg(a + r * ld + c); s1 += __ldg(a + (r + 8) * ld + c);
      s2 += __ldg(a + (r + 16) * ld + c); s3 += __ldg(a + (r + 24) * ld + c);
    }
    for (; r < r1; r += 8) s0 += __ldg(a + r * ld + c);
  }
  sm[ry][threadIdx.x & 31] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ry == 0 && c < C) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    part[(int64_t)blockIdx.y * C + c] = t;
  }
  if (!last_block_of(ticket + blockIdx.x, gridDim.y)) return;
  if (ry == 0 && c < C) {
    float t = 0.f;
    for (int i = 0; i < (int)gridDim.y; ++i) t += __ldcg(part + (int64_t)i * C + c);
    out[c] = t;
  }
}

}  // namespace nvae

using namespace nvae;

int nvae_conv_check(const NvaeConvDesc* d) {
  if (d == nullptr) return NVAE_E_NULLPTR;
  if (d->N <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cin2 < 0 || d->Cout <= 0 || d->R <= 0 || d->S <= 0)
    return NVAE_E_BADSHAPE;
  if (d->stride != 1 && d->stride != 2) return NVAE_E_UNSUPPORTED;
  if (d->y_ld < 0 || d->y_off < 0 || (d->y_ld > 0 && d->y_off + d->Cout > d->y_ld)) return NVAE_E_BADSHAPE;
  if (d->y_ld == 0 && d->y_off != 0) return NVAE_E_BADSHAPE;
  if (d->Ho != (d->H + d->stride - 1) / d->stride || d->Wo != (d->W + d->stride - 1) / d->stride)
    return NVAE_E_BADSHAPE;
  return NVAE_OK;
}

static ConvGeom make_geom(const NvaeConvDesc* d) {
  ConvGeom g;
  g.N = d->N; g.H = d->H; g.W = d->W; g.Cin = d->Cin; g.Cin2 = d->Cin2; g.Ct = d->Cin + d->Cin2; g.Cout = d->Cout;
  g.R = d->R; g.S = d->S; g.stride = d->stride; g.Ho = d->Ho; g.Wo = d->Wo; g.pad_t = d->pad_t; g.pad_l = d->pad_l;
  g.y_ld = d->y_ld > 0 ? d->y_ld : d->Cout;
  g.y_off = d->y_off;
  g.pre_scale = d->pre_scale == 0.f && d->pre_shift == 0.f ? 1.f : d->pre_scale;
  g.pre_shift = d->pre_shift;
  return g;
}

int nvae_colsum(const float* a, int64_t rows, int C, int ld, float* out, void* ws, size_t ws_bytes,
                cudaStream_t stream) {
  int64_t nsplit = ceil_div(rows, 256);
  if (nsplit > 256) nsplit = 256;
  const int64_t rps = ceil_div(rows, nsplit);
  nsplit = ceil_div(rows, rps);
  const int nchunk = (C + 31) / 32;
  if (ws == nullptr || ws_bytes < (size_t)nsplit * C * sizeof(float) + (size_t)nchunk * sizeof(unsigned)) return NVAE_E_WORKSPACE;
  float* part = reinterpret_cast<float*>(ws);
  unsigned* ticket = reinterpret_cast<unsigned*>(part + (size_t)nsplit * C);
  NVAE_CUDA_TRY(cudaMemsetAsync(ticket, 0, (size_t)nchunk * sizeof(unsigned), stream));
  nvae::launch(colsum_kernel, dim3(nchunk, (unsigned)nsplit), 256, 0, stream, a, rows, C, ld, rps, part, ticket, out);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

size_t nvae_colsum_ws_bytes(int C) { return (size_t)256 * C * sizeof(float) + (size_t)((C + 31) / 32) * sizeof(unsigned); }

// Cout == 1 head (postprocess.py:29): y[pix] = b + sum_{tap,c} x[pix+tap][c] * w[tap][c].  Bandwidth-shaped, not
// GEMM-shaped: 8 lanes share a pixel (coalesced 128-byte channel rows), taps come from L1, a 3-step shuffle sums them.
__global__ void __launch_bounds__(256) conv_head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ y, int N,
                                                            int H, int W, int C, int R, int S, int pad_t, int pad_l) {
  nvae::pdl_enter();
  extern __shared__ float wsm[];  // [R*S][C]
  for (int i = threadIdx.x; i < R * S * C; i += blockDim.x) wsm[i] = w[i];
  __syncthreads();
  const int lane8 = threadIdx.x & 7;
  const int64_t npix = (int64_t)N * H * W;
  const float b = bias ? bias[0] : 0.f;
  for (int64_t pix = (int64_t)blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3); pix < npix;
       pix += (int64_t)gridDim.x * (blockDim.x >> 3)) {
    const int wq = (int)(pix % W);
    const int64_t t = pix / W;
    const int hq = (int)(t % H);
    const int64_t n = t / H;
    float acc = 0.f;
    for (int r = 0; r < R; ++r) {
      const int h = hq + r - pad_t;
      if (h < 0 || h >= H) continue;
      for (int s2 = 0; s2 < S; ++s2) {
        const int ww = wq + s2 - pad_l;
        if (ww < 0 || ww >= W) continue;
        const float* xp = x + ((n * H + h) * W + ww) * C;
        const float* wp = wsm + (r * S + s2) * C;
        for (int c = lane8 * 4; c < C; c += 32) {
          const float4 v = ldg4(xp + c);
          const float4 k = *reinterpret_cast<const float4*>(wp + c);
          acc = fmaf(v.x, k.x, acc); acc = fmaf(v.y, k.y, acc); acc = fmaf(v.z, k.z, acc); acc = fmaf(v.w, k.w, acc);
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (lane8 == 0) y[pix] = acc + b;
  }
}

int nvae_conv2d_fwd_simt(const NvaeConvDesc* d, const float* x, const float* x2, const float* w, const float* bias,
                         const float* residual, float* y, cudaStream_t stream) {
  if (d->Cout == 1 && d->stride == 1 && d->Cin2 == 0 && (d->Cin & 3) == 0 && residual == nullptr && d->y_ld <= 1 &&
      d->y_off == 0 && (d->pre_scale == 0.f || d->pre_scale == 1.f) && d->pre_shift == 0.f && d->R * d->S * d->Cin <= 8192 &&
      (reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
    const int64_t npix = (int64_t)d->N * d->H * d->W;
    int64_t grid = ceil_div(npix, 32);
    if (grid > kNumSMs * 16) grid = kNumSMs * 16;
    nvae::launch(conv_head_fwd_kernel, (int)grid, 256, (size_t)d->R * d->S * d->Cin * sizeof(float), stream, 
        x, w, bias, y, d->N, d->H, d->W, d->Cin, d->R, d->S, d->pad_t, d->pad_l);
    NVAE_RETURN_IF_LAUNCH_FAILED();
    return NVAE_OK;
  }
  FwdProb p;
  p.g = make_geom(d);
  p.x = x; p.x2 = x2; p.w = w; p.bias = bias; p.res = residual; p.y = y;
  p.M = (int64_t)d->N * d->Ho * d->Wo;
  p.K = d->R * d->S * p.g.Ct;
  p.Nn = d->Cout;
  dim3 grid((unsigned)ceil_div(p.M, kBM), (unsigned)ceil_div(p.Nn, kBN), 1);
  nvae::launch(simt_conv_kernel<FwdProb>, grid, kSimtThreads, 0, stream, p, p.K, p.K);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

int nvae_conv2d_dgrad_simt(const NvaeConvDesc* d, const float* dy, const float* w, float* dx, float* dx2,
                           int accumulate, cudaStream_t stream) {
  DgradProb p;
  p.g = make_geom(d);
  p.dy = dy; p.w = w; p.dx = dx; p.dx2 = dx2; p.accumulate = accumulate;
  p.M = (int64_t)d->N * d->H * d->W;
  p.K = d->R * d->S * d->Cout;
  p.Nn = p.g.Ct;
  dim3 grid((unsigned)ceil_div(p.M, kBM), (unsigned)ceil_div(p.Nn, kBN), 1);
  nvae::launch(simt_conv_kernel<DgradProb>, grid, kSimtThreads, 0, stream, p, p.K, p.K);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

static void wgrad_split(const NvaeConvDesc* d, int64_t* splits_out, int64_t* kps_out) {
  const int64_t Kpix = (int64_t)d->N * d->Ho * d->Wo;
  const int M = d->R * d->S * (d->Cin + d->Cin2);
  const int64_t tiles = ceil_div(M, kBM) * ceil_div(d->Cout, kBN);
  int64_t splits = ceil_div(4 * kNumSMs, tiles);
  const int64_t max_splits = ceil_div(Kpix, 8 * kBK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int64_t kps = round_up(ceil_div(Kpix, splits), kBK);
  *splits_out = ceil_div(Kpix, kps);
  *kps_out = kps;
}

size_t nvae_conv2d_wgrad_simt_ws_bytes(const NvaeConvDesc* d) {
  int64_t splits, kps;
  wgrad_split(d, &splits, &kps);
  const size_t M = (size_t)d->R * d->S * (d->Cin + d->Cin2);
  return splits > 1 ? (size_t)splits * M * d->Cout * sizeof(float) : 0;
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int64_t n, int splits, float* __restrict__ dw) {
  nvae::pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[(int64_t)k * n + i];
    dw[i] = s;
  }
}

int nvae_conv2d_wgrad_simt(const NvaeConvDesc* d, const float* x, const float* x2, const float* dy, float* dw,
                           void* ws, size_t ws_bytes, cudaStream_t stream) {
  WgradProb p;
  p.g = make_geom(d);
  p.x = x; p.x2 = x2; p.dy = dy; p.dw = dw;
  p.Kpix = (int64_t)d->N * d->Ho * d->Wo;
  p.M = d->R * d->S * p.g.Ct;
  p.Nn = d->Cout;
  int64_t splits, kps;
  wgrad_split(d, &splits, &kps);
  p.split = splits > 1;
  if (p.split) {
    if (ws == nullptr || ws_bytes < nvae_conv2d_wgrad_simt_ws_bytes(d)) return NVAE_E_WORKSPACE;
    p.dw = reinterpret_cast<float*>(ws);
  }
  dim3 grid((unsigned)ceil_div(p.M, kBM), (unsigned)ceil_div(p.Nn, kBN), (unsigned)splits);
  nvae::launch(simt_conv_kernel<WgradProb>, grid, kSimtThreads, 0, stream, p, p.Kpix, kps);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  if (p.split) {
    const int64_t n = (int64_t)p.M * p.Nn;
    int64_t g = ceil_div(n, 256);
    if (g > kNumSMs * 8) g = kNumSMs * 8;
    nvae::launch(wgrad_reduce_kernel, (int)g, 256, 0, stream, p.dw, n, (int)splits, dw);
    NVAE_RETURN_IF_LAUNCH_FAILED();
  }
  return NVAE_OK;
}
