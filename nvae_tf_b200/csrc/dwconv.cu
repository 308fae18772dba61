// Depthwise 5x5 convolution (decoder.py:130) fwd / bwd-data / bwd-filter.  CUDA-core, HBM-bound:
// each CTA stages a zero-haloed [imgs][H+4][W+4..][32ch] tile in shared memory (128-byte channel
// rows -> fully coalesced 128-bit global loads, conflict-free LDS.128), with the BatchNorm-apply +
// swish of decoder.py:141 fused into the staging so the activated tensor never exists in HBM.
#include "common.cuh"

namespace nvae {

constexpr int kDwThreads = 256;
constexpr int kDwCC = 32;  // channels per CTA (8 float4 lanes)
constexpr int kDwTW = 4;   // outputs per thread along W

struct DwGeom {
  int WQ, PW, PH, imgs, ngroups, nchunks;
  size_t smem_tile;  // floats
};

constexpr size_t kDwTileBytes = 44 * 1024;  // haloed tile budget: small tiles -> 3-4 CTAs per SM, loads overlap compute

static DwGeom dw_geom(int N, int H, int W, int C) {
  DwGeom g;
  g.WQ = (W + kDwTW - 1) / kDwTW;
  g.PW = g.WQ * kDwTW + 4;
  g.PH = H + 4;
  const size_t per_img = (size_t)g.PH * g.PW * kDwCC * sizeof(float);
  int imgs = (int)(kDwTileBytes / per_img);
  if (imgs < 1) imgs = 1;
  if (imgs > N) imgs = N;
  g.imgs = imgs;
  g.ngroups = (N + imgs - 1) / imgs;
  g.nchunks = C / kDwCC;
  g.smem_tile = (size_t)imgs * g.PH * g.PW * kDwCC;
  return g;
}

// Stage `nimg` images of the channel chunk into the haloed tile, applying act(x*scale+shift).  The global loads of
// the first batch are issued BEFORE the tile is zero-filled (stores only), so their latency hides behind it; up
// to kDwBatch independent 128-bit loads are in flight per thread.  Ends with the tile complete and synchronised.
constexpr int kDwBatch = 8;
template <bool PROLOGUE>
__device__ __forceinline__ void dw_stage(float* tile, const float* __restrict__ x, const float* __restrict__ stat,
                                         int act, int n0, int nimg, int H, int W, int C, int c0, int PH, int PW) {
  const int c4 = threadIdx.x & 7;
  float4 sc = make_float4(1, 1, 1, 1), sh = make_float4(0, 0, 0, 0);
  if (PROLOGUE && stat != nullptr) {
    sc = ldg4(stat + 2 * C + c0 + c4 * 4);
    sh = ldg4(stat + 3 * C + c0 + c4 * 4);
  }
  const int HW = H * W, npix = nimg * HW;
  constexpr int kStep = kDwThreads >> 3;
  const float* xb = x + (int64_t)n0 * HW * C + c0 + c4 * 4;
  bool zeroed = false;
  for (int p0 = threadIdx.x >> 3; p0 < npix || !zeroed; p0 += kStep * kDwBatch) {
    float4 v[kDwBatch];
#pragma unroll
    for (int j = 0; j < kDwBatch; ++j) {
      const int p = p0 + j * kStep;
      v[j] = p < npix ? ldg4(xb + (int64_t)p * C) : make_float4(0, 0, 0, 0);
    }
    if (!zeroed) {  // uniform across the CTA: every thread runs the first iteration
      const int total4 = nimg * PH * PW * (kDwCC / 4);
      for (int i = threadIdx.x; i < total4; i += kDwThreads) reinterpret_cast<float4*>(tile)[i] = make_float4(0, 0, 0, 0);
      __syncthreads();
      zeroed = true;
    }
#pragma unroll
    for (int j = 0; j < kDwBatch; ++j) {
      const int p = p0 + j * kStep;
      if (p >= npix) continue;
      if (PROLOGUE) {
        v[j].x = act_fwd_rt(fmaf(v[j].x, sc.x, sh.x), act); v[j].y = act_fwd_rt(fmaf(v[j].y, sc.y, sh.y), act);
        v[j].z = act_fwd_rt(fmaf(v[j].z, sc.z, sh.z), act); v[j].w = act_fwd_rt(fmaf(v[j].w, sc.w, sh.w), act);
      }
      const int im = p / HW, q = p - im * HW, h = q / W, w = q - h * W;
      *reinterpret_cast<float4*>(tile + ((size_t)(im * PH + h + 2) * PW + w + 2) * kDwCC + c4 * 4) = v[j];
    }
  }
  __syncthreads();
}

// FLIP=false: y = dwconv(act(bn(x))) + bias.   FLIP=true: da = dwconv_transpose(dy) (no prologue, no bias)
template <bool FLIP>
__global__ void __launch_bounds__(kDwThreads, 2) dwconv5x5_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ stat, int act, int N, int H,
                                                                  int W, int C, const float* __restrict__ wts,
                                                                  const float* __restrict__ bias, float* __restrict__ y,
                                                                  int imgs, int WQ, int PH, int PW) {
  extern __shared__ __align__(16) float smem[];
  float* wsm = smem;             // [25][32]
  float* tile = smem + 25 * kDwCC;
  const int c0 = blockIdx.x * kDwCC, n0 = blockIdx.y * imgs;
  const int nimg = (N - n0) < imgs ? (N - n0) : imgs;
  for (int i = threadIdx.x; i < 25 * kDwCC; i += kDwThreads) {
    const int tap = i / kDwCC, c = i % kDwCC;
    wsm[i] = __ldg(wts + (int64_t)(FLIP ? 24 - tap : tap) * C + c0 + c);
  }
  dw_stage<!FLIP>(tile, x, stat, act, n0, nimg, H, W, C, c0, PH, PW);
  const int c4 = threadIdx.x & 7;
  float4 bv = make_float4(0, 0, 0, 0);
  if (!FLIP && bias != nullptr) bv = ldg4(bias + c0 + c4 * 4);
  const int items = nimg * H * WQ;
  for (int it = threadIdx.x >> 3; it < items; it += kDwThreads >> 3) {
    const int wq = it % WQ, t = it / WQ, h = t % H, im = t / H;
    const int w0 = wq * kDwTW;
    float4 acc[kDwTW];
#pragma unroll
    for (int j = 0; j < kDwTW; ++j) acc[j] = bv;
    const float* trow = tile + ((size_t)(im * PH + h) * PW + w0) * kDwCC + c4 * 4;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      float4 win[kDwTW + 4];
#pragma unroll
      for (int j = 0; j < kDwTW + 4; ++j)
        win[j] = *reinterpret_cast<const float4*>(trow + ((size_t)r * PW + j) * kDwCC);
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        const float4 wv = *reinterpret_cast<const float4*>(wsm + (r * 5 + s) * kDwCC + c4 * 4);
#pragma unroll
        for (int j = 0; j < kDwTW; ++j) {
          acc[j].x = fmaf(win[j + s].x, wv.x, acc[j].x); acc[j].y = fmaf(win[j + s].y, wv.y, acc[j].y);
          acc[j].z = fmaf(win[j + s].z, wv.z, acc[j].z); acc[j].w = fmaf(win[j + s].w, wv.w, acc[j].w);
        }
      }
    }
    float* yo = y + (((int64_t)(n0 + im) * H + h) * W + w0) * C + c0 + c4 * 4;
#pragma unroll
    for (int j = 0; j < kDwTW; ++j)
      if (w0 + j < W) stg4(yo + (int64_t)j * C, acc[j]);
  }
}

// Backward-filter: dw[tap][c] = sum_pix a[pix + tap][c] * dy[pix][c], db[c] = sum dy.  Thread (pixel group pg of 32,
// channel quad c4 of 8) keeps all 25 tap sums + the bias sum in registers and walks its pixels; the 32 pixel groups
// are then combined by two warp shuffles and an 8-warp shared-memory sum, all in fixed order (deterministic).
// partial[g][26][C]: taps 0..24, 25 = sum dy (bias gradient)
__global__ void __launch_bounds__(kDwThreads, 2) dwconv5x5_bwd_filter_kernel(
    const float* __restrict__ x, const float* __restrict__ stat, int act, const float* __restrict__ dy, int N, int H,
    int W, int C, float* __restrict__ partial, int imgs, int PH, int PW) {
  extern __shared__ __align__(16) float smem[];
  float* tile = smem;                                  // haloed activated input
  float* dtile = smem + (size_t)imgs * PH * PW * kDwCC;  // [imgs][H][W][32] dy; reused as the cross-warp buffer
  const int c0 = blockIdx.x * kDwCC, n0 = blockIdx.y * imgs;
  const int nimg = (N - n0) < imgs ? (N - n0) : imgs;
  const int c4 = threadIdx.x & 7, pg = threadIdx.x >> 3;
  const int HW = H * W, npix = nimg * HW;
  {
    constexpr int kStep = kDwThreads >> 3;
    const float* db = dy + (int64_t)n0 * HW * C + c0 + c4 * 4;
    for (int p0 = pg; p0 < npix; p0 += kStep * kDwBatch) {
      float4 v[kDwBatch];
#pragma unroll
      for (int j = 0; j < kDwBatch; ++j) {
        const int p = p0 + j * kStep;
        v[j] = p < npix ? ldg4(db + (int64_t)p * C) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < kDwBatch; ++j) {
        const int p = p0 + j * kStep;
        if (p < npix) *reinterpret_cast<float4*>(dtile + (size_t)p * kDwCC + c4 * 4) = v[j];
      }
    }
  }
  dw_stage<true>(tile, x, stat, act, n0, nimg, H, W, C, c0, PH, PW);  // ends with __syncthreads
  float4 acc[26];
#pragma unroll
  for (int i = 0; i < 26; ++i) acc[i] = make_float4(0, 0, 0, 0);
  for (int p = pg; p < npix; p += kDwThreads >> 3) {
    const int im = p / HW, q = p - im * HW, h = q / W, w = q - h * W;
    const float4 d = *reinterpret_cast<const float4*>(dtile + (size_t)p * kDwCC + c4 * 4);
    const float* ar = tile + ((size_t)(im * PH + h) * PW + w) * kDwCC + c4 * 4;
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        const float4 a = *reinterpret_cast<const float4*>(ar + ((size_t)r * PW + s) * kDwCC);
        float4& t = acc[r * 5 + s];
        t.x = fmaf(a.x, d.x, t.x); t.y = fmaf(a.y, d.y, t.y); t.z = fmaf(a.z, d.z, t.z); t.w = fmaf(a.w, d.w, t.w);
      }
    acc[25].x += d.x; acc[25].y += d.y; acc[25].z += d.z; acc[25].w += d.w;
  }
  // pixel groups of one warp (lanes differing in bits 3,4), then the 8 warps through shared memory
#pragma unroll
  for (int i = 0; i < 26; ++i) {
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
      acc[i].x += __shfl_xor_sync(0xffffffffu, acc[i].x, o); acc[i].y += __shfl_xor_sync(0xffffffffu, acc[i].y, o);
      acc[i].z += __shfl_xor_sync(0xffffffffu, acc[i].z, o); acc[i].w += __shfl_xor_sync(0xffffffffu, acc[i].w, o);
    }
  }
  __syncthreads();  // every thread is done with dtile
  float* red = dtile;  // [8 warps][26][32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 26; ++i) *reinterpret_cast<float4*>(red + ((size_t)warp * 26 + i) * kDwCC + lane * 4) = acc[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 26 * kDwCC; i += kDwThreads) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < kDwThreads / 32; ++wv) s += red[(size_t)wv * 26 * kDwCC + i];
    partial[((int64_t)blockIdx.y * 26 + i / kDwCC) * C + c0 + (i % kDwCC)] = s;
  }
}

__global__ void dwconv5x5_bwd_filter_reduce_kernel(const float* __restrict__ partial, int ngroups, int C,
                                                   float* __restrict__ dw, float* __restrict__ dbias) {
  const int total = 26 * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int g = 0;
    for (; g + 3 < ngroups; g += 4) {  // four loads in flight; fixed association
      s0 += partial[(int64_t)g * total + i]; s1 += partial[(int64_t)(g + 1) * total + i];
      s2 += partial[(int64_t)(g + 2) * total + i]; s3 += partial[(int64_t)(g + 3) * total + i];
    }
    for (; g < ngroups; ++g) s0 += partial[(int64_t)g * total + i];
    const float s = (s0 + s1) + (s2 + s3);
    if (i < 25 * C) dw[i] = s;
    else if (dbias != nullptr) dbias[i - 25 * C] = s;
  }
}

static int dw_check(int N, int H, int W, int C) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C % kDwCC)) return NVAE_E_BADSHAPE;
  DwGeom g = dw_geom(N, H, W, C);
  if ((g.smem_tile * 2 + 26 * kDwCC * 8) * sizeof(float) > 200 * 1024) return NVAE_E_UNSUPPORTED;
  return NVAE_OK;
}

}  // namespace nvae

using namespace nvae;

template <bool FLIP>
static int dw_launch(const float* x, const float* stat, int act, int N, int H, int W, int C, const float* w,
                     const float* bias, float* y, cudaStream_t stream) {
  DwGeom g = dw_geom(N, H, W, C);
  const size_t smem = (g.smem_tile + 25 * kDwCC) * sizeof(float);
  static size_t configured[2] = {0, 0};
  if (smem > configured[FLIP]) {
    NVAE_CUDA_TRY(cudaFuncSetAttribute(dwconv5x5_kernel<FLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured[FLIP] = 200 * 1024;
  }
  dwconv5x5_kernel<FLIP><<<dim3(g.nchunks, g.ngroups), kDwThreads, smem, stream>>>(x, stat, act, N, H, W, C, w, bias, y,
                                                                                   g.imgs, g.WQ, g.PH, g.PW);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_dwconv5x5_fwd(const float* x, const float* stat, int act, int N, int H, int W, int C,
                                  const float* w, const float* bias, float* y, nvae_stream_t stream) {
  int rc = dw_check(N, H, W, C);
  if (rc) return rc;
  if (!x || !w || !y) return NVAE_E_NULLPTR;
  return dw_launch<false>(x, stat, act, N, H, W, C, w, bias, y, stream);
}

extern "C" int nvae_dwconv5x5_bwd_data(const float* dy, int N, int H, int W, int C, const float* w, float* da,
                                       nvae_stream_t stream) {
  int rc = dw_check(N, H, W, C);
  if (rc) return rc;
  if (!dy || !w || !da) return NVAE_E_NULLPTR;
  return dw_launch<true>(dy, nullptr, NVAE_ACT_NONE, N, H, W, C, w, nullptr, da, stream);
}

extern "C" size_t nvae_dwconv5x5_bwd_filter_ws_bytes(int N, int H, int W, int C) {
  if (dw_check(N, H, W, C)) return 0;
  DwGeom g = dw_geom(N, H, W, C);
  return (size_t)g.ngroups * 26 * C * sizeof(float);
}

extern "C" int nvae_dwconv5x5_bwd_filter(const float* x, const float* stat, int act, const float* dy, int N, int H,
                                         int W, int C, float* dw, float* dbias, void* ws, size_t ws_bytes,
                                         nvae_stream_t stream) {
  int rc = dw_check(N, H, W, C);
  if (rc) return rc;
  if (!x || !dy || !dw) return NVAE_E_NULLPTR;
  DwGeom g = dw_geom(N, H, W, C);
  if (ws == nullptr || ws_bytes < (size_t)g.ngroups * 26 * C * sizeof(float)) return NVAE_E_WORKSPACE;
  size_t dfloats = (size_t)g.imgs * H * W * kDwCC;
  if (dfloats < (size_t)(kDwThreads / 32) * 26 * kDwCC) dfloats = (size_t)(kDwThreads / 32) * 26 * kDwCC;  // cross-warp buffer
  const size_t smem = (g.smem_tile + dfloats) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    NVAE_CUDA_TRY(cudaFuncSetAttribute(dwconv5x5_bwd_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       200 * 1024));
    configured = true;
  }
  float* partial = reinterpret_cast<float*>(ws);
  dwconv5x5_bwd_filter_kernel<<<dim3(g.nchunks, g.ngroups), kDwThreads, smem, stream>>>(x, stat, act, dy, N, H, W, C,
                                                                                        partial, g.imgs, g.PH, g.PW);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  dwconv5x5_bwd_filter_reduce_kernel<<<(26 * C + 255) / 256, 256, 0, stream>>>(partial, g.ngroups, C, dw, dbias);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}
