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

constexpr size_t kDwTileBytes = 36 * 1024;  // haloed tile budget: two buffers x 3 CTAs per SM fit the 227 KB of an SM

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

// One-tile-per-CTA variant (the forward direction): y = dwconv(act(bn(x))) + bias with the BatchNorm-apply + activation
// applied in REGISTERS between the global load and the shared-memory store.  (The persistent cp.async kernel below has to
// activate in place in shared memory after the copies land, which measured slower for the forward direction: 45 vs 40 us
// at [144,8,8,768]; for the prologue-free backward-data direction it is the faster one, 31 vs 33 us / 64 vs 74 us.)
// Compute mapping: a warp is the 32 channels of the chunk at one work item, a thread is ONE channel: its 25 taps
// live in registers, shared-memory reads and global stores are 128-byte rows (conflict-free, fully coalesced),
// and each thread produces a 2 x 4 output patch from a 6 x 8 input window (0.24 shared loads per FMA).
template <bool FLIP>
__global__ void __launch_bounds__(kDwThreads, 3) dwconv5x5_tile_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ stat, int act, int N, int H,
                                                                  int W, int C, const float* __restrict__ wts,
                                                                  const float* __restrict__ bias, float* __restrict__ y,
                                                                  int imgs, int WQ, int PH, int PW) {
  nvae::pdl_enter();
  extern __shared__ __align__(16) float smem[];
  float* tile = smem;
  const int c0 = blockIdx.x * kDwCC, n0 = blockIdx.y * imgs;
  const int nimg = (N - n0) < imgs ? (N - n0) : imgs;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float wreg[25];
#pragma unroll
  for (int t = 0; t < 25; ++t) wreg[t] = __ldg(wts + (int64_t)(FLIP ? 24 - t : t) * C + c0 + lane);
  const float bv = (!FLIP && bias != nullptr) ? __ldg(bias + c0 + lane) : 0.f;
  dw_stage<!FLIP>(tile, x, stat, act, n0, nimg, H, W, C, c0, PH, PW);
  const int HP = (H + 1) >> 1;
  const int items = nimg * HP * WQ;
  for (int it = warp; it < items; it += kDwThreads / 32) {
    const int wq = it % WQ, t = it / WQ, hp = t % HP, im = t / HP;
    const int h0 = hp * 2, w0 = wq * kDwTW;
    float acc[2][kDwTW];
#pragma unroll
    for (int j = 0; j < kDwTW; ++j) { acc[0][j] = bv; acc[1][j] = bv; }
    const float* trow = tile + ((size_t)(im * PH + h0) * PW + w0) * kDwCC + lane;
#pragma unroll
    for (int ri = 0; ri < 6; ++ri) {
      if (ri == 5 && h0 + 1 >= H) break;  // the sixth window row only feeds the second output row
      float win[kDwTW + 4];
#pragma unroll
      for (int j = 0; j < kDwTW + 4; ++j) win[j] = trow[((size_t)ri * PW + j) * kDwCC];
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        const int tr = ri - o;
        if (tr < 0 || tr > 4) continue;
#pragma unroll
        for (int s = 0; s < 5; ++s)
#pragma unroll
          for (int j = 0; j < kDwTW; ++j) acc[o][j] = fmaf(win[j + s], wreg[tr * 5 + s], acc[o][j]);
      }
    }
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      if (h0 + o >= H) break;
      float* yo = y + (((int64_t)(n0 + im) * H + h0 + o) * W + w0) * C + c0 + lane;
#pragma unroll
      for (int j = 0; j < kDwTW; ++j)
        if (w0 + j < W) yo[(int64_t)j * C] = acc[o][j];
    }
  }
}

// FLIP=false: y = dwconv(act(bn(x))) + bias.   FLIP=true: da = dwconv_transpose(dy) (no prologue, no bias)
// Persistent + pipelined: a CTA owns one 32-channel chunk (its 25 taps stay in registers) and walks image groups with
// a DOUBLE-BUFFERED haloed tile: while the warps compute group g from one buffer, cp.async (16-byte, L2 -> shared
// memory, no registers) is already filling the other with group g + gridDim.y, so global loads are in flight for the
// whole lifetime of the CTA instead of only during a short staging phase (the one-tile-per-CTA version spent most of
// each CTA's ~10 us on dependent latencies: 22-26 % of the HBM peak).  The halo of both buffers is zeroed once; the
// BatchNorm-apply + activation runs in place on the elements each thread copied itself.
// Compute mapping: a warp is the 32 channels of the chunk at one work item, a thread is ONE channel: shared-memory
// reads and global stores are 128-byte rows (conflict-free, fully coalesced), and each thread produces a 2 x 4 output
// patch from a 6 x 8 input window (0.24 shared loads per FMA).
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The tile geometry is the same for every group a CTA walks, so everything index-shaped is decoded ONCE per CTA into
// registers: the shared-memory offset of each pixel this thread copies (its global offset is just p * C), and the
// (window, output) offsets of each work item this warp computes.  The per-group loops are then copies / FMAs only
// (the integer divisions of the first version were a third of all issued instructions; ncu: issue-bound at 64 %).
constexpr int kDwMaxCopies = 10;  // pixels per thread: ceil(288 / 32) for the largest tile
constexpr int kDwMaxItems = 4;    // work items per warp: ceil(32 / 8)

template <bool FLIP>
__global__ void __launch_bounds__(kDwThreads, 3) dwconv5x5_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ stat, int act, int N, int H,
                                                                  int W, int C, const float* __restrict__ wts,
                                                                  const float* __restrict__ bias, float* __restrict__ y,
                                                                  int imgs, int ngroups, int WQ, int PH, int PW) {
  nvae::pdl_enter();
  extern __shared__ __align__(16) float smem[];
  const int tile_floats = imgs * PH * PW * kDwCC;
  const int c0 = blockIdx.x * kDwCC;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c4 = threadIdx.x & 7;
  const int HW = H * W, HP = (H + 1) >> 1;
  // zero both buffers once: the halo stays zero, the interiors are overwritten by every group's copies
  for (int i = threadIdx.x; i < 2 * tile_floats / 4; i += kDwThreads) reinterpret_cast<float4*>(smem)[i] = make_float4(0, 0, 0, 0);
  float wreg[25];
#pragma unroll
  for (int t = 0; t < 25; ++t) wreg[t] = __ldg(wts + (int64_t)(FLIP ? 24 - t : t) * C + c0 + lane);
  const float bv = (!FLIP && bias != nullptr) ? __ldg(bias + c0 + lane) : 0.f;
  float4 sc = make_float4(1, 1, 1, 1), sh = make_float4(0, 0, 0, 0);
  const bool prologue = !FLIP && stat != nullptr;
  if (prologue) {
    sc = ldg4(stat + 2 * C + c0 + c4 * 4);
    sh = ldg4(stat + 3 * C + c0 + c4 * 4);
  }
  const bool activate = prologue || (!FLIP && act != NVAE_ACT_NONE);
  // copy k of this thread: pixel p = (tid >> 3) + 32 k of the group -> shared-memory float offset (-1: none)
  int soff[kDwMaxCopies];
#pragma unroll
  for (int k = 0; k < kDwMaxCopies; ++k) {
    const int p = (threadIdx.x >> 3) + k * (kDwThreads >> 3);
    const int im = p / HW, q = p - im * HW, h = q / W, w = q - h * W;
    soff[k] = p < imgs * HW ? ((im * PH + h + 2) * PW + w + 2) * kDwCC + c4 * 4 : -1;
  }
  // work item m of this warp: it = warp + 8 m -> window offset in the tile, output offset in the group, image, h0, w0
  // (kept in shared memory: the item loop is not unrolled -- four copies of its 200-FMA body would not fit the
  // instruction cache -- and a register table cannot be indexed dynamically)
  __shared__ int itab[kDwThreads / 32][kDwMaxItems][3];
  if (lane < kDwMaxItems) {
    const int it = warp + lane * (kDwThreads / 32);
    const int wq = it % WQ, t = it / WQ, hp = t % HP, im = t / HP;
    const int h0 = hp * 2, w0 = wq * kDwTW;
    itab[warp][lane][0] = ((im * PH + h0) * PW + w0) * kDwCC;
    itab[warp][lane][1] = ((im * H + h0) * W + w0) * C;
    itab[warp][lane][2] = it < imgs * HP * WQ ? (im << 16) | (h0 << 8) | w0 : -1;
  }
  __syncthreads();
  const float* xc = x + c0 + c4 * 4;
  float* yc = y + c0 + lane;
  const int p0 = threadIdx.x >> 3;
  auto issue = [&](float* tile, int g) {
    const int n0 = g * imgs;
    const int npix = ((N - n0) < imgs ? (N - n0) : imgs) * HW;
    const float* xb = xc + (int64_t)n0 * HW * C;
#pragma unroll
    for (int k = 0; k < kDwMaxCopies; ++k) {
      const int p = p0 + k * (kDwThreads >> 3);
      if (soff[k] >= 0 && p < npix) cp_async16(tile + soff[k], xb + (int64_t)p * C);
    }
  };
  int g = blockIdx.y, buf = 0;
  if (g < ngroups) issue(smem, g);
  cp_async_commit();
  for (; g < ngroups; g += gridDim.y, buf ^= 1) {
    float* tile = smem + (size_t)buf * tile_floats;
    const int n0 = g * imgs;
    const int nimg = (N - n0) < imgs ? (N - n0) : imgs;
    // next group -> the other buffer (its readers finished at the barrier that ended the last pass)
    if (g + (int)gridDim.y < ngroups) issue(smem + (size_t)(buf ^ 1) * tile_floats, g + gridDim.y);
    cp_async_commit();
    cp_async_wait<1>();  // this group's copies have landed (only the newest commit may be pending)
    if (activate) {      // act(x * scale + shift) in place on exactly the elements this thread copied itself
      const int npix = nimg * HW;
#pragma unroll
      for (int k = 0; k < kDwMaxCopies; ++k) {
        if (soff[k] >= 0 && p0 + k * (kDwThreads >> 3) < npix) {
          float4* e = reinterpret_cast<float4*>(tile + soff[k]);
          float4 v = *e;
          v.x = act_fwd_rt(fmaf(v.x, sc.x, sh.x), act); v.y = act_fwd_rt(fmaf(v.y, sc.y, sh.y), act);
          v.z = act_fwd_rt(fmaf(v.z, sc.z, sh.z), act); v.w = act_fwd_rt(fmaf(v.w, sc.w, sh.w), act);
          *e = v;
        }
      }
    }
    __syncthreads();
    float* yg = yc + (int64_t)n0 * HW * C;
#pragma unroll 1
    for (int m = 0; m < kDwMaxItems; ++m) {
      const int code = itab[warp][m][2];
      if (code < 0 || (code >> 16) >= nimg) break;  // items are ordered by image
      const int h0 = (code >> 8) & 0xff, w0 = code & 0xff;
      float acc[2][kDwTW];
#pragma unroll
      for (int j = 0; j < kDwTW; ++j) { acc[0][j] = bv; acc[1][j] = bv; }
      const float* trow = tile + itab[warp][m][0] + lane;
#pragma unroll
      for (int ri = 0; ri < 6; ++ri) {
        if (ri == 5 && h0 + 1 >= H) break;  // the sixth window row only feeds the second output row
        float win[kDwTW + 4];
#pragma unroll
        for (int j = 0; j < kDwTW + 4; ++j) win[j] = trow[(ri * PW + j) * kDwCC];
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const int tr = ri - o;
          if (tr < 0 || tr > 4) continue;
#pragma unroll
          for (int s = 0; s < 5; ++s)
#pragma unroll
            for (int j = 0; j < kDwTW; ++j) acc[o][j] = fmaf(win[j + s], wreg[tr * 5 + s], acc[o][j]);
        }
      }
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        if (h0 + o >= H) break;
        float* yo = yg + itab[warp][m][1] + (int64_t)o * W * C;
#pragma unroll
        for (int j = 0; j < kDwTW; ++j)
          if (w0 + j < W) yo[(int64_t)j * C] = acc[o][j];
      }
    }
    __syncthreads();  // every warp is done with this buffer before the next pass refills it
  }
  cp_async_wait<0>();
}

// Backward-filter: dw[tap][c] = sum_pix a[pix + tap][c] * dy[pix][c], db[c] = sum dy.  A thread is one channel (a
// warp = the chunk's 32 channels at one image row): it keeps the dy row and one input row in registers and
// accumulates all 25 taps + the bias sum (0.3 shared loads per FMA).  The 8 warps are then combined through
// shared memory in warp order (deterministic).  WT = row width the registers are sized for (>= W).
// partial[g][26][C]: taps 0..24, 25 = sum dy (bias gradient)
template <int WT>
__global__ void __launch_bounds__(kDwThreads, 2) dwconv5x5_bwd_filter_kernel(
    const float* __restrict__ x, const float* __restrict__ stat, int act, const float* __restrict__ dy, int N, int H,
    int W, int C, float* __restrict__ partial, int imgs, int PH, int PW) {
  nvae::pdl_enter();
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
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[26];
#pragma unroll
  for (int i = 0; i < 26; ++i) acc[i] = 0.f;
  for (int it = warp; it < nimg * H; it += kDwThreads / 32) {
    const int im = it / H, h = it - im * H;
    float dyr[WT];
    const float* dr = dtile + ((size_t)(im * H + h) * W) * kDwCC + lane;
#pragma unroll
    for (int j = 0; j < WT; ++j) {
      dyr[j] = j < W ? dr[(size_t)j * kDwCC] : 0.f;
      acc[25] += dyr[j];
    }
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      float xr[WT + 4];
      const float* ar = tile + ((size_t)(im * PH + h + r) * PW) * kDwCC + lane;
#pragma unroll
      for (int j = 0; j < WT + 4; ++j) xr[j] = j < W + 4 ? ar[(size_t)j * kDwCC] : 0.f;
#pragma unroll
      for (int s = 0; s < 5; ++s)
#pragma unroll
        for (int j = 0; j < WT; ++j) acc[r * 5 + s] = fmaf(xr[j + s], dyr[j], acc[r * 5 + s]);
    }
  }
  __syncthreads();  // every thread is done with dtile
  float* red = dtile;  // [8 warps][26][32]
#pragma unroll
  for (int i = 0; i < 26; ++i) red[((size_t)warp * 26 + i) * kDwCC + lane] = acc[i];
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
  nvae::pdl_enter();
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
  if (g.imgs * H * W > kDwMaxCopies * (kDwThreads >> 3) || g.imgs * ((H + 1) / 2) * g.WQ > kDwMaxItems * (kDwThreads / 32) ||
      H > 255 || W > 255)
    return NVAE_E_UNSUPPORTED;  // per-thread copy / per-warp item tables of dwconv5x5_kernel
  return NVAE_OK;
}

}  // namespace nvae

using namespace nvae;

template <bool FLIP>
static int dw_launch(const float* x, const float* stat, int act, int N, int H, int W, int C, const float* w,
                     const float* bias, float* y, cudaStream_t stream) {
  DwGeom g = dw_geom(N, H, W, C);
  if (!FLIP) {  // forward: one tile per CTA, activation in registers
    const size_t smem1 = g.smem_tile * sizeof(float);
    static bool configured1 = false;
    if (!configured1) {
      NVAE_CUDA_TRY(cudaFuncSetAttribute(dwconv5x5_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured1 = true;
    }
    nvae::launch(dwconv5x5_tile_kernel<false>, dim3(g.nchunks, g.ngroups), kDwThreads, smem1, stream, x, stat, act, N, H, W, C, w,
                 bias, y, g.imgs, g.WQ, g.PH, g.PW);
    NVAE_RETURN_IF_LAUNCH_FAILED();
    return NVAE_OK;
  }
  const size_t smem = 2 * g.smem_tile * sizeof(float);  // double buffer
  static size_t configured[2] = {0, 0};
  if (smem > configured[FLIP]) {
    NVAE_CUDA_TRY(cudaFuncSetAttribute(dwconv5x5_kernel<FLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured[FLIP] = 200 * 1024;
  }
  // persistent: ~3 CTAs per SM in total, each walking its chunk's image groups
  int gy = (int)ceil_div(3 * kNumSMs, g.nchunks);
  if (gy > g.ngroups) gy = g.ngroups;
  nvae::launch(dwconv5x5_kernel<FLIP>, dim3(g.nchunks, gy), kDwThreads, smem, stream, x, stat, act, N, H, W, C, w, bias, y,
               g.imgs, g.ngroups, g.WQ, g.PH, g.PW);
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
  float* partial = reinterpret_cast<float*>(ws);
  const dim3 grid(g.nchunks, g.ngroups);
#define NVAE_DW_FILTER(WT_)                                                                                           \
  do {                                                                                                                \
    static bool configured = false;                                                                                   \
    if (!configured) {                                                                                                \
      NVAE_CUDA_TRY(cudaFuncSetAttribute(dwconv5x5_bwd_filter_kernel<WT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         200 * 1024));                                                                \
      configured = true;                                                                                              \
    }                                                                                                                 \
    nvae::launch(dwconv5x5_bwd_filter_kernel<WT_>, grid, kDwThreads, smem, stream, x, stat, act, dy, N, H, W, C, partial, g.imgs, \
                                                                         g.PH, g.PW);                                 \
  } while (0)
  if (W <= 4) NVAE_DW_FILTER(4);
  else if (W <= 8) NVAE_DW_FILTER(8);
  else if (W <= 16) NVAE_DW_FILTER(16);
  else return NVAE_E_UNSUPPORTED;
#undef NVAE_DW_FILTER
  NVAE_RETURN_IF_LAUNCH_FAILED();
  nvae::launch(dwconv5x5_bwd_filter_reduce_kernel, (26 * C + 255) / 256, 256, 0, stream, partial, g.ngroups, C, dw, dbias);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}
