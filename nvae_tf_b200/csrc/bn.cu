// BatchNorm statistics / apply+activation / backward -- bandwidth-bound, 128-bit NHWC access.
// Reference semantics: layers.BatchNormalization(momentum=0.05, epsilon=1e-5) followed by
// activations.swish / ELU (common.py:148,166-167; encoder.py:91-104; decoder.py:125-145).
#include <stdlib.h>

#include "common.cuh"

namespace nvae {

constexpr int kBnWarps = 8;
constexpr int kBnThreads = kBnWarps * 32;
constexpr int kBnMaxSplit = 1024;

struct BnGeom {
  int C4, LC, RW, nchunk, nsplit;
  int64_t rows_per_split;
};

static BnGeom bn_geom(int64_t rows, int C) {
  BnGeom g;
  g.C4 = C / 4;
  int lc = 1;
  while (lc < g.C4 && lc < 32) lc <<= 1;
  // channel lanes per CTA: a chunk width that does not divide C/4 leaves idle lanes in the last chunk (C = 192: 48 float4
  // lanes as 32 + 16 -> a quarter of all threads idle); take the widest power of two >= 8 (128-byte row segments) with
  // the fewest wasted lanes
  if (lc > 8) {
    int best = lc, best_waste = (int)(ceil_div(g.C4, lc) * lc) - g.C4;
    for (int w = lc >> 1; w >= 8; w >>= 1) {
      const int waste = (int)(ceil_div(g.C4, w) * w) - g.C4;
      if (waste < best_waste) { best = w; best_waste = waste; }
    }
    lc = best;
  }
  g.LC = lc;
  g.RW = 32 / lc;
  g.nchunk = (int)ceil_div(g.C4, lc);
  const int64_t rows_iter = (int64_t)kBnWarps * g.RW;
  int64_t want = ceil_div(4 * kNumSMs, g.nchunk);            // ~4 CTAs per SM in total
  int64_t cap = ceil_div(rows, rows_iter * 4);               // >= 4 iterations of work per CTA
  int64_t ns = want < cap ? want : cap;
  if (ns < 1) ns = 1;
  if (ns > kBnMaxSplit) ns = kBnMaxSplit;
  g.rows_per_split = round_up(ceil_div(rows, ns), rows_iter);
  g.nsplit = (int)ceil_div(rows, g.rows_per_split);
  return g;
}

__device__ __forceinline__ void f4_acc(float4& s, float4& q, const float4& v, const float4& k) {
  float a = v.x - k.x, b = v.y - k.y, c = v.z - k.z, d = v.w - k.w;
  s.x += a; s.y += b; s.z += c; s.w += d;
  q.x += a * a; q.y += b * b; q.z += c * c; q.w += d * d;
}

// Cross-lane (row sub-groups of a warp) and cross-warp reduction of two float4 accumulators.
// Result valid in threads with warp==0 && lane<LC.  sm: [kBnWarps][32][8] floats.
__device__ __forceinline__ void bn_block_reduce(float4& s, float4& q, int LC, float (*sm)[32][8]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = LC; o < 32; o <<= 1) {
    s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
    s.z += __shfl_xor_sync(0xffffffffu, s.z, o); s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
    q.x += __shfl_xor_sync(0xffffffffu, q.x, o); q.y += __shfl_xor_sync(0xffffffffu, q.y, o);
    q.z += __shfl_xor_sync(0xffffffffu, q.z, o); q.w += __shfl_xor_sync(0xffffffffu, q.w, o);
  }
  float* d = sm[warp][lane];
  d[0] = s.x; d[1] = s.y; d[2] = s.z; d[3] = s.w; d[4] = q.x; d[5] = q.y; d[6] = q.z; d[7] = q.w;
  __syncthreads();
  if (warp == 0) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int w = 0; w < kBnWarps; ++w)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += sm[w][lane][i];
    s = make_float4(acc[0], acc[1], acc[2], acc[3]);
    q = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// part: [nsplit][2][C] doubles (mean, M2) using a per-CTA pivot (shifted-data algorithm)
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const float* __restrict__ x, int64_t rows, int C,
                                                              int64_t rows_per_split, int LC,
                                                              double* __restrict__ part) {
  nvae::pdl_enter();
  __shared__ float sm[kBnWarps][32][8];
  const int C4 = C >> 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int RW = 32 / LC, rsub = lane / LC, cl = lane % LC;
  const int c4 = blockIdx.x * LC + cl;
  const bool cvalid = c4 < C4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r1 = r0 + rows_per_split < rows ? r0 + rows_per_split : rows;
  float4 s = make_float4(0, 0, 0, 0), q = s, k = s;
  if (cvalid) {
    const float* xc = x + (int64_t)c4 * 4;
    k = ldg4(xc + r0 * C);
    const int64_t step = (int64_t)kBnWarps * RW;
    int64_t r = r0 + warp * RW + rsub;
    for (; r + 3 * step < r1; r += 4 * step) {
      float4 v0 = ldg4(xc + r * C), v1 = ldg4(xc + (r + step) * C), v2 = ldg4(xc + (r + 2 * step) * C),
             v3 = ldg4(xc + (r + 3 * step) * C);
      f4_acc(s, q, v0, k); f4_acc(s, q, v1, k); f4_acc(s, q, v2, k); f4_acc(s, q, v3, k);
    }
    for (; r < r1; r += step) f4_acc(s, q, ldg4(xc + r * C), k);
  }
  bn_block_reduce(s, q, LC, sm);
  if (warp == 0 && lane < LC && cvalid) {
    const double n = (double)(r1 - r0);
    double* pm = part + ((int64_t)blockIdx.y * 2) * C + c4 * 4;
    double* pq = pm + C;
    const float sv[4] = {s.x, s.y, s.z, s.w}, qv[4] = {q.x, q.y, q.z, q.w}, kv[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pm[i] = (double)kv[i] + (double)sv[i] / n;
      pq[i] = (double)qv[i] - (double)sv[i] * (double)sv[i] / n;
    }
  }
}

// Combines the per-CTA (mean, M2) partials: one warp per channel, lane l takes splits l, l+32, ...; two passes of
// plain sums (mean = sum n_s*m_s / n;  M2 = sum q_s + n_s*(m_s - mean)^2) and a fixed shuffle tree -- no serial
// chain of divisions, deterministic.
__global__ void __launch_bounds__(256) bn_finalize_kernel(const double* __restrict__ part, int nsplit, int64_t rows,
                                                          int64_t rows_per_split, int C,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ mm,
                                                          float* __restrict__ mv, int training, float momentum,
                                                          float eps, float* __restrict__ stat) {
  nvae::pdl_enter();
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= C) return;
  float mean, var;
  if (training) {
    const double n = (double)rows;
    const double n_last = (double)(rows - (int64_t)(nsplit - 1) * rows_per_split), n_full = (double)rows_per_split;
    // ONE pass over the partials (they are fp64, so the textbook sum n_s*m_s^2 - n*mu^2 has bits to spare; the shifted
    // per-CTA pivots keep m_s small): three plain sums, one shuffle tree each -- half the dependent L2 round trips of
    // the mean-then-M2 formulation
    double a0 = 0, a1 = 0, a2 = 0;
#pragma unroll 4
    for (int s = lane; s < nsplit; s += 32) {
      const double ns = s == nsplit - 1 ? n_last : n_full;
      const double m = part[((int64_t)s * 2) * C + c], q = part[((int64_t)s * 2 + 1) * C + c];
      a0 += ns * m;
      a1 += q;
      a2 += ns * m * m;
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    const double mu = a0 / n;
    double M2 = a1 + a2 - n * mu * mu;
    if (M2 < 0) M2 = 0;
    if (lane != 0) return;
    mean = (float)mu;
    var = (float)(M2 / n);
    if (mm != nullptr) {
      const double unbiased = n > 1 ? M2 / (n - 1) : M2 / n;
      mm[c] = mm[c] * momentum + mean * (1.f - momentum);
      mv[c] = mv[c] * momentum + (float)unbiased * (1.f - momentum);
    }
  } else {
    if (lane != 0) return;
    mean = mm[c];
    var = mv[c];
  }
  const float invstd = rsqrtf(var + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float scale = g * invstd;
  stat[c] = mean;
  stat[C + c] = invstd;
  stat[2 * C + c] = scale;
  stat[3 * C + c] = b - mean * scale;
}

// ---- forward apply + activation (+ nearest x2 upsample, + TF32 rounding) ---------------------
template <int ACT>
__global__ void bn_act_fwd_kernel(const float* __restrict__ x, int64_t n4, int C4, const float* __restrict__ stat,
                                  int up_h, int up_w, int round_mode, float* __restrict__ out) {
  nvae::pdl_enter();
  const int C = C4 * 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    float4 v = ldg4(x + i * 4);
    if (stat != nullptr) {
      const float4 sc = ldg4(stat + 2 * C + c4 * 4), sh = ldg4(stat + 3 * C + c4 * 4);
      v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    }
    v.x = act_fwd<ACT>(v.x); v.y = act_fwd<ACT>(v.y); v.z = act_fwd<ACT>(v.z); v.w = act_fwd<ACT>(v.w);
    if (round_mode) v = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
    if (up_h == 0) {
      stg4(out + i * 4, v);
    } else {
      const int64_t row = i / C4;
      const int w = (int)(row % up_w);
      const int64_t t = row / up_w;
      const int h = (int)(t % up_h);
      const int64_t n = t / up_h;
      const int64_t W2 = 2 * (int64_t)up_w;
      const int64_t o00 = (((n * 2 * up_h + 2 * h) * W2) + 2 * w) * C + c4 * 4;
      stg4(out + o00, v); stg4(out + o00 + C, v); stg4(out + o00 + W2 * C, v); stg4(out + o00 + W2 * C + C, v);
    }
  }
}

// ---- backward ------------------------------------------------------------------------------
__device__ __forceinline__ float4 load_dout(const float* __restrict__ dout, int64_t row, int C, int c4, int up_h,
                                            int up_w) {
  if (up_h == 0) return ldg4(dout + row * C + c4 * 4);
  const int w = (int)(row % up_w);
  const int64_t t = row / up_w;
  const int h = (int)(t % up_h);
  const int64_t n = t / up_h;
  const int64_t W2 = 2 * (int64_t)up_w;
  const float* p = dout + (((n * 2 * up_h + 2 * h) * W2) + 2 * w) * C + c4 * 4;
  const float4 a = ldg4(p), b = ldg4(p + C), c = ldg4(p + W2 * C), d = ldg4(p + W2 * C + C);
  return make_float4((a.x + b.x) + (c.x + d.x), (a.y + b.y) + (c.y + d.y), (a.z + b.z) + (c.z + d.z),
                     (a.w + b.w) + (c.w + d.w));
}

template <int ACT>
__device__ __forceinline__ void g_and_xhat(const float4& dv, const float4& xv, const float4& mean,
                                           const float4& invstd, const float4& sc, const float4& sh, float4& g,
                                           float4& xh) {
  xh = make_float4((xv.x - mean.x) * invstd.x, (xv.y - mean.y) * invstd.y, (xv.z - mean.z) * invstd.z,
                   (xv.w - mean.w) * invstd.w);
  g = make_float4(dv.x * act_grad<ACT>(fmaf(xv.x, sc.x, sh.x)), dv.y * act_grad<ACT>(fmaf(xv.y, sc.y, sh.y)),
                  dv.z * act_grad<ACT>(fmaf(xv.z, sc.z, sh.z)), dv.w * act_grad<ACT>(fmaf(xv.w, sc.w, sh.w)));
}

// part: [nsplit][2][C] doubles: sum g, sum g*xhat
template <int ACT>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_reduce_kernel(const float* __restrict__ dout,
                                                                   const float* __restrict__ x, int64_t rows, int C,
                                                                   int64_t rows_per_split, int LC,
                                                                   const float* __restrict__ stat, int up_h, int up_w,
                                                                   double* __restrict__ part) {
  nvae::pdl_enter();
  __shared__ float sm[kBnWarps][32][8];
  const int C4 = C >> 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int RW = 32 / LC, rsub = lane / LC, cl = lane % LC;
  const int c4 = blockIdx.x * LC + cl;
  const bool cvalid = c4 < C4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r1 = r0 + rows_per_split < rows ? r0 + rows_per_split : rows;
  float4 s = make_float4(0, 0, 0, 0), q = s;
  if (cvalid) {
    const float4 mean = ldg4(stat + c4 * 4), invstd = ldg4(stat + C + c4 * 4), sc = ldg4(stat + 2 * C + c4 * 4),
                 sh = ldg4(stat + 3 * C + c4 * 4);
    const int64_t step = (int64_t)kBnWarps * RW;
#pragma unroll 2
    for (int64_t r = r0 + warp * RW + rsub; r < r1; r += step) {
      const float4 dv = load_dout(dout, r, C, c4, up_h, up_w);
      const float4 xv = ldg4(x + r * C + c4 * 4);
      float4 g, xh;
      g_and_xhat<ACT>(dv, xv, mean, invstd, sc, sh, g, xh);
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
      q.x += g.x * xh.x; q.y += g.y * xh.y; q.z += g.z * xh.z; q.w += g.w * xh.w;
    }
  }
  bn_block_reduce(s, q, LC, sm);
  if (warp == 0 && lane < LC && cvalid) {
    double* ps = part + ((int64_t)blockIdx.y * 2) * C + c4 * 4;
    double* pq = ps + C;
    ps[0] = s.x; ps[1] = s.y; ps[2] = s.z; ps[3] = s.w;
    pq[0] = q.x; pq[1] = q.y; pq[2] = q.z; pq[3] = q.w;
  }
}

// bstat: [2][C] floats: mean(g), mean(g*xhat)
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const double* __restrict__ part, int nsplit,
                                                              int64_t rows, int C, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, float* __restrict__ bstat) {
  nvae::pdl_enter();
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);  // one warp per channel
  if (c >= C) return;
  double sg = 0, sq = 0;
#pragma unroll 4
  for (int s = lane; s < nsplit; s += 32) {
    sg += part[((int64_t)s * 2) * C + c];
    sq += part[((int64_t)s * 2 + 1) * C + c];
  }
  sg = warp_sum(sg);
  sq = warp_sum(sq);
  if (lane != 0) return;
  if (dgamma) dgamma[c] = (float)sq;
  if (dbeta) dbeta[c] = (float)sg;
  bstat[c] = (float)(sg / (double)rows);
  bstat[C + c] = (float)(sq / (double)rows);
}

template <int ACT>
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ x, int64_t rows, int C4,
                                    const float* __restrict__ stat, const float* __restrict__ bstat, int up_h,
                                    int up_w, const float* __restrict__ dres, float res_scale, int accumulate,
                                    float* __restrict__ dx) {
  nvae::pdl_enter();
  const int C = C4 * 4;
  const int64_t n4 = rows * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const int64_t row = i / C4;
    const float4 dv = load_dout(dout, row, C, c4, up_h, up_w);
    const float4 xv = ldg4(x + i * 4);
    float4 r;
    if (stat != nullptr) {
      const float4 mean = ldg4(stat + c4 * 4), invstd = ldg4(stat + C + c4 * 4), sc = ldg4(stat + 2 * C + c4 * 4),
                   sh = ldg4(stat + 3 * C + c4 * 4);
      float4 g, xh;
      g_and_xhat<ACT>(dv, xv, mean, invstd, sc, sh, g, xh);
      if (bstat != nullptr) {
        const float4 mg = ldg4(bstat + c4 * 4), mq = ldg4(bstat + C + c4 * 4);
        r = make_float4(sc.x * (g.x - mg.x - xh.x * mq.x), sc.y * (g.y - mg.y - xh.y * mq.y),
                        sc.z * (g.z - mg.z - xh.z * mq.z), sc.w * (g.w - mg.w - xh.w * mq.w));
      } else {
        r = make_float4(sc.x * g.x, sc.y * g.y, sc.z * g.z, sc.w * g.w);
      }
    } else {
      r = make_float4(dv.x * act_grad<ACT>(xv.x), dv.y * act_grad<ACT>(xv.y), dv.z * act_grad<ACT>(xv.z),
                      dv.w * act_grad<ACT>(xv.w));
    }
    if (dres != nullptr) {
      const float4 d = ldg4(dres + i * 4);
      r.x = fmaf(res_scale, d.x, r.x); r.y = fmaf(res_scale, d.y, r.y); r.z = fmaf(res_scale, d.z, r.z);
      r.w = fmaf(res_scale, d.w, r.w);
    }
    if (accumulate) {
      const float4 o = *reinterpret_cast<const float4*>(dx + i * 4);
      r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
    }
    stg4(dx + i * 4, r);
  }
}

// ---- cluster-fused BatchNorm (tensors that live in L2) -----------------------------------------
// One launch does what bn_stats + bn_finalize (+ bn_act_fwd), or bn_bwd_reduce + bn_bwd_finalize + bn_bwd_apply,
// do in three.  A channel chunk (4*LC channels) is owned by one thread-block CLUSTER of S CTAs that split the
// rows; each CTA reduces its rows, publishes its per-channel partials in its own shared memory, and after a cluster
// barrier every CTA combines the S partials through distributed shared memory in rank order (fixed association,
// fp64 -> bit-repeatable and identical in all CTAs of the cluster).  The apply pass then re-reads exactly the
// lines this CTA streamed a moment ago (L1/L2 hits).  Used when the tensor is at most kBnFusedMaxBytes; larger
// tensors keep the split kernels, whose grids fill all SMs.
constexpr int64_t kBnFusedMaxBytes = 40ll << 20;
constexpr int kBnMaxCluster = 8;  // portable cluster size

struct BnFusedGeom {
  int LC, RW, nchunk, S;
  int64_t rows_per_cta;
  bool ok;
};

static BnFusedGeom bn_fused_geom(int64_t rows, int C, bool backward) {
  BnFusedGeom g;
  const int C4 = C / 4;
  int lc = 1;
  while (lc < C4 && lc < 32) lc <<= 1;
  int S = kBnMaxCluster;
  while (S > 1 && rows < (int64_t)S * kBnWarps * (32 / lc) * 2) S >>= 1;
  while (lc > 2 && ceil_div(C4, lc) * S < 2 * kNumSMs) lc >>= 1;  // narrower chunks -> more clusters (>= 2 CTAs per SM)
  g.LC = lc;
  g.RW = 32 / lc;
  g.nchunk = (int)ceil_div(C4, lc);
  g.S = S;
  g.rows_per_cta = round_up(ceil_div(rows, S), (int64_t)kBnWarps * g.RW);
  const int64_t bytes = rows * (int64_t)C * 4;
  static const bool enabled = [] {  // NVAE_BN_FUSED=0: always the split kernels (A/B measurements)
    const char* e = getenv("NVAE_BN_FUSED");
    return !(e != nullptr && e[0] == '0');
  }();
  // measured inside a CUDA graph (profiles/r01i): forward (two passes over x) wins up to ~30 MB while a row segment
  // is >= 64 bytes; backward (five passes, two operands) only for the smallest tensors -- its few fat CTAs lose to
  // the split kernels' 4 CTAs per SM as soon as the tensor is more than a few MB
  g.ok = enabled && (backward ? bytes <= (3ll << 20) : bytes <= kBnFusedMaxBytes && (g.LC >= 4 || bytes <= (3ll << 20)));
  return g;
}

template <int ACT>
__global__ void __launch_bounds__(kBnThreads) bn_fwd_fused_kernel(
    const float* __restrict__ x, int64_t rows, int C, int64_t rows_per_cta, int LC, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* __restrict__ mm, float* __restrict__ mv, float momentum, float eps,
    float* __restrict__ stat, int up_h, int up_w, int round_mode, float* __restrict__ out) {
  pdl_enter();
  __shared__ float sm[kBnWarps][32][8];
  __shared__ double part[2][128];  // this CTA's per-channel (mean, M2) over its rows
  __shared__ float sstat[2][128];  // scale, shift of the chunk
  const int C4 = C >> 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int RW = 32 / LC, rsub = lane / LC, cl = lane % LC;
  const int c4 = blockIdx.x * LC + cl;
  const bool cvalid = c4 < C4;
  const int rank = blockIdx.y, S = gridDim.y;  // cluster = (1, S, 1)
  const int64_t r0 = (int64_t)rank * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  const int64_t step = (int64_t)kBnWarps * RW;
  float4 s = make_float4(0, 0, 0, 0), q = s, k = s;
  const float* xc = x + (int64_t)c4 * 4;
  if (cvalid && r0 < r1) {
    k = ldg4(xc + r0 * C);
    int64_t r = r0 + warp * RW + rsub;
    for (; r + 7 * step < r1; r += 8 * step) {  // eight 128-bit loads in flight per thread
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ldg4(xc + (r + j * step) * C);
#pragma unroll
      for (int j = 0; j < 8; ++j) f4_acc(s, q, v[j], k);
    }
    for (; r < r1; r += step) f4_acc(s, q, ldg4(xc + r * C), k);
  }
  bn_block_reduce(s, q, LC, sm);
  if (warp == 0 && lane < LC) {
    const double n = r1 > r0 ? (double)(r1 - r0) : 1.0;
    const float sv[4] = {s.x, s.y, s.z, s.w}, qv[4] = {q.x, q.y, q.z, q.w}, kv[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      part[0][lane * 4 + i] = (double)kv[i] + (double)sv[i] / n;
      part[1][lane * 4 + i] = (double)qv[i] - (double)sv[i] * (double)sv[i] / n;
    }
  }
  cluster_barrier();
  if (threadIdx.x < 4 * LC) {
    const int c = blockIdx.x * 4 * LC + threadIdx.x;
    if (c < C) {
      double m_s[kBnMaxCluster], q_s[kBnMaxCluster], n_s[kBnMaxCluster];
#pragma unroll
      for (int j = 0; j < kBnMaxCluster; ++j) {
        if (j < S) {
          m_s[j] = dsmem_ld_f64(&part[0][threadIdx.x], j);
          q_s[j] = dsmem_ld_f64(&part[1][threadIdx.x], j);
          const int64_t a = (int64_t)j * rows_per_cta, b = a + rows_per_cta < rows ? a + rows_per_cta : rows;
          n_s[j] = b > a ? (double)(b - a) : 0.0;
        } else {
          m_s[j] = q_s[j] = n_s[j] = 0.0;
        }
      }
      const double n = (double)rows;
      double mu = 0;
#pragma unroll
      for (int j = 0; j < kBnMaxCluster; ++j) mu += n_s[j] * m_s[j];
      mu /= n;
      double M2 = 0;
#pragma unroll
      for (int j = 0; j < kBnMaxCluster; ++j) M2 += n_s[j] > 0 ? q_s[j] + n_s[j] * (m_s[j] - mu) * (m_s[j] - mu) : 0.0;
      const float mean = (float)mu, var = (float)(M2 / n);
      const float invstd = rsqrtf(var + eps);
      const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
      const float scale = g * invstd, shift = b - mean * scale;
      sstat[0][threadIdx.x] = scale;
      sstat[1][threadIdx.x] = shift;
      if (rank == 0) {
        if (mm != nullptr) {
          const double unbiased = n > 1 ? M2 / (n - 1) : M2 / n;
          mm[c] = mm[c] * momentum + mean * (1.f - momentum);
          mv[c] = mv[c] * momentum + (float)unbiased * (1.f - momentum);
        }
        stat[c] = mean;
        stat[C + c] = invstd;
        stat[2 * C + c] = scale;
        stat[3 * C + c] = shift;
      }
    }
  }
  __syncthreads();
  if (out != nullptr && cvalid) {
    const float4 sc = *reinterpret_cast<const float4*>(&sstat[0][cl * 4]);
    const float4 sh = *reinterpret_cast<const float4*>(&sstat[1][cl * 4]);
    const int64_t W2 = 2 * (int64_t)up_w;
    for (int64_t rb = r0 + warp * RW + rsub; rb < r1; rb += 8 * step) {
      float4 vv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t r = rb + j * step;
        vv[j] = r < r1 ? ldg4(xc + r * C) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t r = rb + j * step;
        if (r >= r1) break;
        float4 v = vv[j];
        v.x = act_fwd<ACT>(fmaf(v.x, sc.x, sh.x)); v.y = act_fwd<ACT>(fmaf(v.y, sc.y, sh.y));
        v.z = act_fwd<ACT>(fmaf(v.z, sc.z, sh.z)); v.w = act_fwd<ACT>(fmaf(v.w, sc.w, sh.w));
        if (round_mode) v = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
        if (up_h == 0) {
          stg4(out + r * C + c4 * 4, v);
        } else {
          const int w = (int)(r % up_w);
          const int64_t t = r / up_w;
          const int h = (int)(t % up_h);
          const int64_t n = t / up_h;
          const int64_t o00 = (((n * 2 * up_h + 2 * h) * W2) + 2 * w) * C + c4 * 4;
          stg4(out + o00, v); stg4(out + o00 + C, v); stg4(out + o00 + W2 * C, v); stg4(out + o00 + W2 * C + C, v);
        }
      }
    }
  }
  cluster_barrier();  // no CTA leaves while a peer may still read its partials
}

template <int ACT>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_fused_kernel(
    const float* __restrict__ dout, const float* __restrict__ x, int64_t rows, int C, int64_t rows_per_cta, int LC,
    const float* __restrict__ stat, int up_h, int up_w, const float* __restrict__ dres, float res_scale,
    int accumulate, float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_enter();
  __shared__ float sm[kBnWarps][32][8];
  __shared__ double part[2][128];  // this CTA's per-channel sum g, sum g*xhat
  __shared__ float sb[2][128];     // mean(g), mean(g*xhat) of the chunk
  const int C4 = C >> 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int RW = 32 / LC, rsub = lane / LC, cl = lane % LC;
  const int c4 = blockIdx.x * LC + cl;
  const bool cvalid = c4 < C4;
  const int rank = blockIdx.y, S = gridDim.y;
  const int64_t r0 = (int64_t)rank * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  const int64_t step = (int64_t)kBnWarps * RW;
  float4 s = make_float4(0, 0, 0, 0), q = s;
  float4 mean = s, invstd = s, sc = s, sh = s;
  if (cvalid) {
    mean = ldg4(stat + c4 * 4); invstd = ldg4(stat + C + c4 * 4); sc = ldg4(stat + 2 * C + c4 * 4);
    sh = ldg4(stat + 3 * C + c4 * 4);
    for (int64_t rb = r0 + warp * RW + rsub; rb < r1; rb += 4 * step) {  // 2 x 4 loads in flight per thread
      float4 dv[4], xv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t r = rb + j * step;
        if (r < r1) {
          dv[j] = load_dout(dout, r, C, c4, up_h, up_w);
          xv[j] = ldg4(x + r * C + c4 * 4);
        } else {
          dv[j] = xv[j] = make_float4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (rb + j * step >= r1) break;
        float4 g, xh;
        g_and_xhat<ACT>(dv[j], xv[j], mean, invstd, sc, sh, g, xh);
        s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
        q.x += g.x * xh.x; q.y += g.y * xh.y; q.z += g.z * xh.z; q.w += g.w * xh.w;
      }
    }
  }
  bn_block_reduce(s, q, LC, sm);
  if (warp == 0 && lane < LC) {
    part[0][lane * 4 + 0] = s.x; part[0][lane * 4 + 1] = s.y; part[0][lane * 4 + 2] = s.z; part[0][lane * 4 + 3] = s.w;
    part[1][lane * 4 + 0] = q.x; part[1][lane * 4 + 1] = q.y; part[1][lane * 4 + 2] = q.z; part[1][lane * 4 + 3] = q.w;
  }
  cluster_barrier();
  if (threadIdx.x < 4 * LC) {
    const int c = blockIdx.x * 4 * LC + threadIdx.x;
    if (c < C) {
      double sg = 0, sq = 0;
#pragma unroll
      for (int j = 0; j < kBnMaxCluster; ++j)
        if (j < S) {
          sg += dsmem_ld_f64(&part[0][threadIdx.x], j);
          sq += dsmem_ld_f64(&part[1][threadIdx.x], j);
        }
      sb[0][threadIdx.x] = (float)(sg / (double)rows);
      sb[1][threadIdx.x] = (float)(sq / (double)rows);
      if (rank == 0) {
        if (dgamma) dgamma[c] = (float)sq;
        if (dbeta) dbeta[c] = (float)sg;
      }
    }
  }
  __syncthreads();
  if (cvalid) {
    const float4 mg = *reinterpret_cast<const float4*>(&sb[0][cl * 4]);
    const float4 mq = *reinterpret_cast<const float4*>(&sb[1][cl * 4]);
    for (int64_t rb = r0 + warp * RW + rsub; rb < r1; rb += 4 * step) {
      float4 dvv[4], xvv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t r = rb + j * step;
        if (r < r1) {
          dvv[j] = load_dout(dout, r, C, c4, up_h, up_w);
          xvv[j] = ldg4(x + r * C + c4 * 4);
        } else {
          dvv[j] = xvv[j] = make_float4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
      const int64_t r = rb + j * step;
      if (r >= r1) break;
      const int64_t o = r * C + c4 * 4;
      float4 g, xh;
      g_and_xhat<ACT>(dvv[j], xvv[j], mean, invstd, sc, sh, g, xh);
      float4 rr = make_float4(sc.x * (g.x - mg.x - xh.x * mq.x), sc.y * (g.y - mg.y - xh.y * mq.y),
                              sc.z * (g.z - mg.z - xh.z * mq.z), sc.w * (g.w - mg.w - xh.w * mq.w));
      if (dres != nullptr) {
        const float4 d = ldg4(dres + o);
        rr.x = fmaf(res_scale, d.x, rr.x); rr.y = fmaf(res_scale, d.y, rr.y); rr.z = fmaf(res_scale, d.z, rr.z);
        rr.w = fmaf(res_scale, d.w, rr.w);
      }
      if (accumulate) {
        const float4 old = *reinterpret_cast<const float4*>(dx + o);
        rr.x += old.x; rr.y += old.y; rr.z += old.z; rr.w += old.w;
      }
      stg4(dx + o, rr);
      }
    }
  }
  cluster_barrier();
}

static int ew_grid(int64_t n, int threads) {
  int64_t b = ceil_div(n, threads);
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace nvae

using namespace nvae;

extern "C" size_t nvae_bn_ws_bytes(int64_t rows, int C) {
  (void)rows;
  return (size_t)kBnMaxSplit * 2 * C * sizeof(double) + (size_t)2 * C * sizeof(float);
}

extern "C" int nvae_bn_stats(const float* x, int64_t rows, int C, const float* gamma, const float* beta,
                             float* moving_mean, float* moving_var, int training, float momentum, float eps,
                             float* stat, void* ws, size_t ws_bytes, nvae_stream_t stream) {
  if (C <= 0 || (C & 3) || rows <= 0) return NVAE_E_BADSHAPE;
  if (stat == nullptr || (training && x == nullptr)) return NVAE_E_NULLPTR;
  if (!training && (moving_mean == nullptr || moving_var == nullptr)) return NVAE_E_NULLPTR;
  if (training) {
    const BnFusedGeom f = bn_fused_geom(rows, C, false);
    if (f.ok) {  // stats + finalize in one cluster launch
      launch_cluster(bn_fwd_fused_kernel<NVAE_ACT_NONE>, dim3(f.nchunk, f.S), kBnThreads, 0, stream, dim3(1, f.S, 1), x,
                     rows, C, f.rows_per_cta, f.LC, gamma, beta, moving_mean, moving_var, momentum, eps, stat, 0, 0, 0,
                     (float*)nullptr);
      NVAE_RETURN_IF_LAUNCH_FAILED();
      return NVAE_OK;
    }
  }
  BnGeom g = bn_geom(rows, C);
  double* part = reinterpret_cast<double*>(ws);
  if (training) {
    if (ws == nullptr || ws_bytes < (size_t)g.nsplit * 2 * C * sizeof(double)) return NVAE_E_WORKSPACE;
    nvae::launch(bn_stats_kernel, dim3(g.nchunk, g.nsplit), kBnThreads, 0, stream, x, rows, C, g.rows_per_split, g.LC, part);
    NVAE_RETURN_IF_LAUNCH_FAILED();
  }
  nvae::launch(bn_finalize_kernel, (C + 7) / 8, 256, 0, stream, part, g.nsplit, rows, g.rows_per_split, C, gamma, beta,
                                                           moving_mean, moving_var, training, momentum, eps, stat);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

template <int ACT>
static void bn_fwd_fused_launch(const BnFusedGeom& f, const float* x, int64_t rows, int C, const float* gamma,
                                const float* beta, float* mm, float* mv, float momentum, float eps, float* stat, int up_h,
                                int up_w, int round_tf32, float* out, cudaStream_t stream) {
  launch_cluster(bn_fwd_fused_kernel<ACT>, dim3(f.nchunk, f.S), kBnThreads, 0, stream, dim3(1, f.S, 1), x, rows, C,
                 f.rows_per_cta, f.LC, gamma, beta, mm, mv, momentum, eps, stat, up_h, up_w, round_tf32, out);
}

extern "C" int nvae_bn_fwd(const float* x, int64_t rows, int C, const float* gamma, const float* beta,
                           float* moving_mean, float* moving_var, int training, float momentum, float eps, float* stat,
                           int act, int up_h, int up_w, int round_tf32, float* out, void* ws, size_t ws_bytes,
                           nvae_stream_t stream) {
  if (C <= 0 || (C & 3) || rows <= 0) return NVAE_E_BADSHAPE;
  if (x == nullptr || out == nullptr || stat == nullptr) return NVAE_E_NULLPTR;
  if ((up_h == 0) != (up_w == 0)) return NVAE_E_BADSHAPE;
  if (up_h && rows % ((int64_t)up_h * up_w)) return NVAE_E_BADSHAPE;
  if (act != NVAE_ACT_NONE && act != NVAE_ACT_SWISH && act != NVAE_ACT_ELU) return NVAE_E_UNSUPPORTED;
  const BnFusedGeom f = bn_fused_geom(rows, C, false);
  if (training && f.ok) {
    if (act == NVAE_ACT_NONE)
      bn_fwd_fused_launch<NVAE_ACT_NONE>(f, x, rows, C, gamma, beta, moving_mean, moving_var, momentum, eps, stat, up_h,
                                         up_w, round_tf32, out, stream);
    else if (act == NVAE_ACT_SWISH)
      bn_fwd_fused_launch<NVAE_ACT_SWISH>(f, x, rows, C, gamma, beta, moving_mean, moving_var, momentum, eps, stat, up_h,
                                          up_w, round_tf32, out, stream);
    else
      bn_fwd_fused_launch<NVAE_ACT_ELU>(f, x, rows, C, gamma, beta, moving_mean, moving_var, momentum, eps, stat, up_h,
                                        up_w, round_tf32, out, stream);
    NVAE_RETURN_IF_LAUNCH_FAILED();
    return NVAE_OK;
  }
  int rc = nvae_bn_stats(x, rows, C, gamma, beta, moving_mean, moving_var, training, momentum, eps, stat, ws, ws_bytes,
                         stream);
  if (rc) return rc;
  return nvae_bn_act_fwd(x, rows, C, stat, act, up_h, up_w, round_tf32, out, stream);
}

extern "C" int nvae_bn_act_fwd(const float* x, int64_t rows, int C, const float* stat, int act, int up_h, int up_w,
                               int round_tf32, float* out, nvae_stream_t stream) {
  if (C <= 0 || (C & 3) || rows <= 0) return NVAE_E_BADSHAPE;
  if (x == nullptr || out == nullptr) return NVAE_E_NULLPTR;
  if ((up_h == 0) != (up_w == 0)) return NVAE_E_BADSHAPE;
  if (up_h && rows % ((int64_t)up_h * up_w)) return NVAE_E_BADSHAPE;
  const int64_t n4 = rows * (C / 4);
  const int grid = ew_grid(n4, 256);
  switch (act) {
    case NVAE_ACT_NONE:
      nvae::launch(bn_act_fwd_kernel<NVAE_ACT_NONE>, grid, 256, 0, stream, x, n4, C / 4, stat, up_h, up_w, round_tf32, out);
      break;
    case NVAE_ACT_SWISH:
      nvae::launch(bn_act_fwd_kernel<NVAE_ACT_SWISH>, grid, 256, 0, stream, x, n4, C / 4, stat, up_h, up_w, round_tf32, out);
      break;
    case NVAE_ACT_ELU:
      nvae::launch(bn_act_fwd_kernel<NVAE_ACT_ELU>, grid, 256, 0, stream, x, n4, C / 4, stat, up_h, up_w, round_tf32, out);
      break;
    default:
      return NVAE_E_UNSUPPORTED;
  }
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

template <int ACT>
static int bn_act_bwd_impl(const float* dout, const float* x, int64_t rows, int C, const float* stat, int up_h,
                           int up_w, int training, const float* dres, float res_scale, int accumulate, float* dx,
                           float* dgamma, float* dbeta, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (stat != nullptr && training && dx != nullptr) {
    const BnFusedGeom f = bn_fused_geom(rows, C, true);
    if (f.ok) {  // reduce + finalize + apply in one cluster launch
      launch_cluster(bn_bwd_fused_kernel<ACT>, dim3(f.nchunk, f.S), kBnThreads, 0, stream, dim3(1, f.S, 1), dout, x, rows,
                     C, f.rows_per_cta, f.LC, stat, up_h, up_w, dres, res_scale, accumulate, dx, dgamma, dbeta);
      NVAE_RETURN_IF_LAUNCH_FAILED();
      return NVAE_OK;
    }
  }
  float* bstat = nullptr;
  if (stat != nullptr && (training || dgamma != nullptr || dbeta != nullptr)) {
    BnGeom g = bn_geom(rows, C);
    const size_t need = (size_t)g.nsplit * 2 * C * sizeof(double) + (size_t)2 * C * sizeof(float);
    if (ws == nullptr || ws_bytes < need) return NVAE_E_WORKSPACE;
    double* part = reinterpret_cast<double*>(ws);
    float* bs = reinterpret_cast<float*>(part + (size_t)g.nsplit * 2 * C);
    nvae::launch(bn_bwd_reduce_kernel<ACT>, dim3(g.nchunk, g.nsplit), kBnThreads, 0, stream, dout, x, rows, C, g.rows_per_split,
                                                                                  g.LC, stat, up_h, up_w, part);
    NVAE_RETURN_IF_LAUNCH_FAILED();
    nvae::launch(bn_bwd_finalize_kernel, (C + 7) / 8, 256, 0, stream, part, g.nsplit, rows, C, dgamma, dbeta, bs);
    NVAE_RETURN_IF_LAUNCH_FAILED();
    if (training) bstat = bs;
  }
  if (dx != nullptr) {
    nvae::launch(bn_bwd_apply_kernel<ACT>, ew_grid(rows * (C / 4), 256), 256, 0, stream, dout, x, rows, C / 4, stat, bstat, up_h,
                                                                              up_w, dres, res_scale, accumulate, dx);
    NVAE_RETURN_IF_LAUNCH_FAILED();
  }
  return NVAE_OK;
}

extern "C" int nvae_bn_act_bwd(const float* dout, const float* x, int64_t rows, int C, const float* stat, int act,
                               int up_h, int up_w, int training, const float* dres, float res_scale, int accumulate,
                               float* dx, float* dgamma, float* dbeta, void* ws, size_t ws_bytes,
                               nvae_stream_t stream) {
  if (C <= 0 || (C & 3) || rows <= 0) return NVAE_E_BADSHAPE;
  if (dout == nullptr || x == nullptr) return NVAE_E_NULLPTR;
  if ((up_h == 0) != (up_w == 0)) return NVAE_E_BADSHAPE;
  switch (act) {
    case NVAE_ACT_NONE:
      return bn_act_bwd_impl<NVAE_ACT_NONE>(dout, x, rows, C, stat, up_h, up_w, training, dres, res_scale, accumulate,
                                            dx, dgamma, dbeta, ws, ws_bytes, stream);
    case NVAE_ACT_SWISH:
      return bn_act_bwd_impl<NVAE_ACT_SWISH>(dout, x, rows, C, stat, up_h, up_w, training, dres, res_scale, accumulate,
                                             dx, dgamma, dbeta, ws, ws_bytes, stream);
    case NVAE_ACT_ELU:
      return bn_act_bwd_impl<NVAE_ACT_ELU>(dout, x, rows, C, stat, up_h, up_w, training, dres, res_scale, accumulate,
                                           dx, dgamma, dbeta, ws, ws_bytes, stream);
    default:
      return NVAE_E_UNSUPPORTED;
  }
}
