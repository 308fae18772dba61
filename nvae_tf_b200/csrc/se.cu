// Squeeze-and-Excitation gate + residual merge (common.py:129-142 fused with encoder.py:107,
// decoder.py:147, preprocess.py:107, postprocess.py:58).  Bandwidth-bound: one pass over t for the
// global pool (+ the two tiny dense layers), one pass (cache hits) for the gated residual -- one cluster launch.
#include "common.cuh"

namespace nvae {

constexpr int kSeThreads = 256;
constexpr int kSeMaxC = 1024;
constexpr int kSeMaxHid = 64;

// ---- cluster versions: S CTAs per sample -------------------------------------------------------------------------
// One CTA per sample leaves the machine almost empty (144 CTAs of 256 threads on 148 SMs, each walking its sample with
// four loads in flight: 15-29 % of the HBM rate on tensors beyond L2).  Here a sample is owned by a thread-block CLUSTER
// of S <= 8 CTAs that split its rows: each CTA pools its rows, the per-channel partial sums are exchanged through
// distributed shared memory and added in rank order (fixed association: every CTA of the cluster gets the same bits),
// every CTA then computes the (tiny) dense -> relu -> dense -> sigmoid chain redundantly and applies the gate to its own
// rows, which it streamed a moment ago (L1/L2 hits).  Forward and backward are ONE launch each at any tensor size.
__device__ __forceinline__ int se_cluster_rows(int HW, int S, int rank, int* r0) {
  const int rps = (HW + S - 1) / S;
  *r0 = rank * rps;
  const int r1 = *r0 + rps < HW ? *r0 + rps : HW;
  return r1 > *r0 ? r1 : *r0;
}

__global__ void __launch_bounds__(kSeThreads) se_fwd_cluster_kernel(
    const float* __restrict__ t, const float* __restrict__ stat, int HW, int C, int hid, const float* __restrict__ w1,
    const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
    float* __restrict__ pooled, float* __restrict__ hidden, float* __restrict__ gate, const float* __restrict__ xres,
    float alpha, float beta, float* __restrict__ y, int S) {
  nvae::pdl_enter();
  __shared__ float part[kSeThreads * 4];
  __shared__ __align__(16) float cpart[kSeMaxC];  // this CTA's per-channel partial sums (read by the cluster)
  __shared__ __align__(16) float spool[kSeMaxC];
  __shared__ float shid[kSeMaxHid];
  const int b = blockIdx.x / S, rank = blockIdx.x % S, C4 = C >> 2, tid = threadIdx.x;
  const int G = kSeThreads / C4, grp = tid / C4, c4 = tid % C4;
  int r0;
  const int r1 = se_cluster_rows(HW, S, rank, &r0);
  const float* tb = t + (int64_t)b * HW * C;
  float4 acc = make_float4(0, 0, 0, 0);
  if (grp < G) {
    int r = r0 + grp;
    for (; r + 7 * G < r1; r += 8 * G) {  // eight 128-bit loads in flight per thread
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ldg4(tb + (int64_t)(r + j * G) * C + c4 * 4);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
    }
    for (; r < r1; r += G) {
      const float4 v = ldg4(tb + (int64_t)r * C + c4 * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(&part[tid * 4]) = acc;
  }
  __syncthreads();
  for (int c = tid; c < C; c += kSeThreads) {
    float s = 0.f;
    for (int g = 0; g < G; ++g) s += part[(g * C4 + (c >> 2)) * 4 + (c & 3)];
    cpart[c] = s;
  }
  cluster_barrier();
  for (int c = tid; c < C; c += kSeThreads) {
    float s = 0.f;
    for (int j = 0; j < S; ++j) s += dsmem_ld_f32(&cpart[c], j);
    s *= 1.f / (float)HW;
    if (stat != nullptr) s = fmaf(s, stat[2 * C + c], stat[3 * C + c]);
    spool[c] = s;
    if (rank == 0) pooled[(int64_t)b * C + c] = s;
  }
  cluster_barrier();  // every CTA is done reading its peers' partials: from here on CTAs are independent
  const int lane = tid & 31, warp = tid >> 5;
  for (int j = warp; j < hid; j += kSeThreads / 32) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(spool[c], __ldg(w1 + (int64_t)c * hid + j), s);
    s = warp_sum(s);
    if (lane == 0) {
      s = fmaxf(s + b1[j], 0.f);
      shid[j] = s;
      if (rank == 0) hidden[(int64_t)b * hid + j] = s;
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += kSeThreads) {
    float s = b2[c];
    for (int j = 0; j < hid; ++j) s = fmaf(shid[j], __ldg(w2 + (int64_t)j * C + c), s);
    s = 1.f / (1.f + expf(-s));
    if (rank == 0) gate[(int64_t)b * C + c] = s;
    spool[c] = s;  // (the pooled values are no longer needed)
  }
  __syncthreads();
  // gated residual merge of this CTA's rows: y = alpha*xres + beta*t'*gate
  if (grp >= G) return;
  float4 sc = make_float4(1, 1, 1, 1), sh = make_float4(0, 0, 0, 0);
  if (stat != nullptr) {
    sc = ldg4(stat + 2 * C + c4 * 4);
    sh = ldg4(stat + 3 * C + c4 * 4);
  }
  const float4 g = *reinterpret_cast<const float4*>(&spool[c4 * 4]);
  const float bx = beta * g.x, by = beta * g.y, bz = beta * g.z, bw = beta * g.w;
  const float* xb = xres + (int64_t)b * HW * C;
  float* yb = y + (int64_t)b * HW * C;
  for (int rb = r0 + grp; rb < r1; rb += 4 * G) {
    float4 v[4], xr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = rb + j * G;
      if (r < r1) {
        v[j] = ldg4(tb + (int64_t)r * C + c4 * 4);
        xr[j] = ldg4(xb + (int64_t)r * C + c4 * 4);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = rb + j * G;
      if (r >= r1) break;
      float4 o;
      o.x = fmaf(alpha, xr[j].x, bx * fmaf(v[j].x, sc.x, sh.x)); o.y = fmaf(alpha, xr[j].y, by * fmaf(v[j].y, sc.y, sh.y));
      o.z = fmaf(alpha, xr[j].z, bz * fmaf(v[j].z, sc.z, sh.z)); o.w = fmaf(alpha, xr[j].w, bw * fmaf(v[j].w, sc.w, sh.w));
      stg4(yb + (int64_t)r * C + c4 * 4, o);
    }
  }
}

// Backward: r[c] = beta*sum_hw dy*t' (cluster reduction), chain through sigmoid / dense2 / relu / dense1 in every CTA,
// rank 0 writes dz2 [B,C], dh [B,hid], dpool [B,C] for the dense-layer gradients; then on this CTA's rows
// dt' = beta*dy*gate + dpool/HW and dxres (+)= alpha*dy.
__global__ void __launch_bounds__(kSeThreads) se_bwd_cluster_kernel(
    const float* __restrict__ dy, const float* __restrict__ t, const float* __restrict__ stat, int HW, int C, int hid,
    const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ hidden,
    const float* __restrict__ gate, float beta, float* __restrict__ dz2, float* __restrict__ dh,
    float* __restrict__ dpool, float alpha, float* __restrict__ dt, float* __restrict__ dxres, int dxres_accumulate,
    int S) {
  nvae::pdl_enter();
  __shared__ float part[kSeThreads * 4];
  __shared__ __align__(16) float cpart[kSeMaxC];
  __shared__ __align__(16) float sdz[kSeMaxC];
  __shared__ float sdh[kSeMaxHid];
  const int b = blockIdx.x / S, rank = blockIdx.x % S, C4 = C >> 2, tid = threadIdx.x;
  const int G = kSeThreads / C4, grp = tid / C4, c4 = tid % C4;
  int r0;
  const int r1 = se_cluster_rows(HW, S, rank, &r0);
  const float* tb = t + (int64_t)b * HW * C;
  const float* db = dy + (int64_t)b * HW * C;
  float4 acc = make_float4(0, 0, 0, 0);
  if (grp < G) {
    float4 sc = make_float4(1, 1, 1, 1), sh = make_float4(0, 0, 0, 0);
    if (stat != nullptr) {
      sc = ldg4(stat + 2 * C + c4 * 4);
      sh = ldg4(stat + 3 * C + c4 * 4);
    }
    int r = r0 + grp;
    for (; r + 3 * G < r1; r += 4 * G) {  // eight 128-bit loads in flight per thread
      float4 v[4], d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = ldg4(tb + (int64_t)(r + j * G) * C + c4 * 4);
        d[j] = ldg4(db + (int64_t)(r + j * G) * C + c4 * 4);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc.x = fmaf(d[j].x, fmaf(v[j].x, sc.x, sh.x), acc.x); acc.y = fmaf(d[j].y, fmaf(v[j].y, sc.y, sh.y), acc.y);
        acc.z = fmaf(d[j].z, fmaf(v[j].z, sc.z, sh.z), acc.z); acc.w = fmaf(d[j].w, fmaf(v[j].w, sc.w, sh.w), acc.w);
      }
    }
    for (; r < r1; r += G) {
      const float4 v = ldg4(tb + (int64_t)r * C + c4 * 4), d = ldg4(db + (int64_t)r * C + c4 * 4);
      acc.x = fmaf(d.x, fmaf(v.x, sc.x, sh.x), acc.x); acc.y = fmaf(d.y, fmaf(v.y, sc.y, sh.y), acc.y);
      acc.z = fmaf(d.z, fmaf(v.z, sc.z, sh.z), acc.z); acc.w = fmaf(d.w, fmaf(v.w, sc.w, sh.w), acc.w);
    }
    *reinterpret_cast<float4*>(&part[tid * 4]) = acc;
  }
  __syncthreads();
  for (int c = tid; c < C; c += kSeThreads) {
    float s = 0.f;
    for (int g = 0; g < G; ++g) s += part[(g * C4 + (c >> 2)) * 4 + (c & 3)];
    cpart[c] = s;
  }
  cluster_barrier();
  for (int c = tid; c < C; c += kSeThreads) {
    float s = 0.f;
    for (int j = 0; j < S; ++j) s += dsmem_ld_f32(&cpart[c], j);
    const float gt = gate[(int64_t)b * C + c];
    const float d = beta * s * gt * (1.f - gt);
    sdz[c] = d;
    if (rank == 0) dz2[(int64_t)b * C + c] = d;
  }
  cluster_barrier();  // peers are done with this CTA's partials
  const int lane = tid & 31, warp = tid >> 5;
  for (int j = warp; j < hid; j += kSeThreads / 32) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(sdz[c], __ldg(w2 + (int64_t)j * C + c), s);
    s = warp_sum(s);
    if (lane == 0) {
      s = hidden[(int64_t)b * hid + j] > 0.f ? s : 0.f;
      sdh[j] = s;
      if (rank == 0) dh[(int64_t)b * hid + j] = s;
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += kSeThreads) {
    float s = 0.f;
    for (int j = 0; j < hid; ++j) s = fmaf(sdh[j], __ldg(w1 + (int64_t)c * hid + j), s);
    if (rank == 0) dpool[(int64_t)b * C + c] = s;
    sdz[c] = s;  // (every thread rewrites only the entries it read above... after the barrier below all are dpool)
  }
  __syncthreads();
  if (grp >= G) return;
  const float inv_hw = 1.f / (float)HW;
  const float4 g = ldg4(gate + (int64_t)b * C + c4 * 4);
  const float4 p = *reinterpret_cast<const float4*>(&sdz[c4 * 4]);
  const float gx = beta * g.x, gy = beta * g.y, gz = beta * g.z, gw = beta * g.w;
  const float px = p.x * inv_hw, py = p.y * inv_hw, pz = p.z * inv_hw, pw = p.w * inv_hw;
  float* dtb = dt + (int64_t)b * HW * C;
  float* dxb = dxres != nullptr ? dxres + (int64_t)b * HW * C : nullptr;
  for (int rb = r0 + grp; rb < r1; rb += 4 * G) {
    float4 d[4], e[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = rb + j * G;
      if (r < r1) {
        d[j] = ldg4(db + (int64_t)r * C + c4 * 4);
        if (dxb != nullptr && dxres_accumulate) e[j] = *reinterpret_cast<const float4*>(dxb + (int64_t)r * C + c4 * 4);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = rb + j * G;
      if (r >= r1) break;
      float4 o;
      o.x = fmaf(d[j].x, gx, px); o.y = fmaf(d[j].y, gy, py); o.z = fmaf(d[j].z, gz, pz); o.w = fmaf(d[j].w, gw, pw);
      stg4(dtb + (int64_t)r * C + c4 * 4, o);
      if (dxb != nullptr) {
        float4 q = make_float4(alpha * d[j].x, alpha * d[j].y, alpha * d[j].z, alpha * d[j].w);
        if (dxres_accumulate) { q.x += e[j].x; q.y += e[j].y; q.z += e[j].z; q.w += e[j].w; }
        stg4(dxb + (int64_t)r * C + c4 * 4, q);
      }
    }
  }
}

// dense-layer gradients: fixed-order sums over the batch (deterministic).  Each output is owned by 8 adjacent lanes
// that take interleaved batch elements (b = sub, sub + 8, ...), six pairs of loads in flight per lane, and are combined
// by a fixed xor-shuffle tree.  (One thread per output walking the whole batch was a chain of B/4 dependent L2 round
// trips: 20 us for 8 500 dot products of length 144.)
__device__ __forceinline__ float se_dot_batch8(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                               int B, int sub) {
  float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int i = sub;
  for (; i + 40 < B; i += 48) {
    float av[6], bv[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      av[j] = __ldg(a + (int64_t)(i + 8 * j) * lda);
      bv[j] = b ? __ldg(b + (int64_t)(i + 8 * j) * ldb) : 1.f;
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) s[j] = fmaf(av[j], bv[j], s[j]);
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {  // at most five elements are left per lane
    const int ii = i + 8 * j;
    if (ii < B) s[j] = fmaf(__ldg(a + (int64_t)ii * lda), b ? __ldg(b + (int64_t)ii * ldb) : 1.f, s[j]);
  }
  float r = ((s[0] + s[1]) + (s[2] + s[3])) + (s[4] + s[5]);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 4);
  return r;
}

__global__ void se_bwd_weights_kernel(const float* __restrict__ pooled, const float* __restrict__ hidden,
                                      const float* __restrict__ dz2, const float* __restrict__ dh, int B, int C,
                                      int hid, float* __restrict__ dw1, float* __restrict__ db1,
                                      float* __restrict__ dw2, float* __restrict__ db2) {
  nvae::pdl_enter();
  const int n1 = C * hid, total = 2 * n1 + C + hid;
  const int sub = threadIdx.x & 7;
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const bool valid = i0 < total;
  const int i = valid ? i0 : total - 1;  // whole warps reach the shuffles
  const float* a;
  const float* b = nullptr;
  float* out;
  int lda, ldb = 0;
  if (i < n1) {  // dw2[j][c]
    const int j = i / C, c = i % C;
    a = hidden + j; lda = hid; b = dz2 + c; ldb = C; out = dw2 + i;
  } else if (i < 2 * n1) {  // dw1[c][j]
    const int k = i - n1, c = k / hid, j = k % hid;
    a = pooled + c; lda = C; b = dh + j; ldb = hid; out = dw1 + k;
  } else if (i < 2 * n1 + C) {
    const int c = i - 2 * n1;
    a = dz2 + c; lda = C; out = db2 + c;
  } else {
    const int j = i - 2 * n1 - C;
    a = dh + j; lda = hid; out = db1 + j;
  }
  const float r = se_dot_batch8(a, lda, b, ldb, B, sub);
  if (valid && sub == 0) *out = r;
}

}  // namespace nvae

using namespace nvae;

// CTAs per sample: a power of two <= 8 (portable cluster size) with at least one pass of rows per CTA, doubled until the
// grid holds ~4 CTAs per SM
static int se_cluster_size(int B, int HW, int C) {
  const int G = kSeThreads / (C / 4);
  int S = 1;
  while (S < 8 && (int64_t)B * S < 4 * kNumSMs && 2 * S * G <= HW) S *= 2;
  return S;
}

static int se_check(int B, int HW, int C, int hid) {
  if (B <= 0 || HW <= 0 || C <= 0 || (C & 3) || C > kSeMaxC || hid <= 0 || hid > kSeMaxHid) return NVAE_E_BADSHAPE;
  if (C / 4 > kSeThreads) return NVAE_E_BADSHAPE;
  return NVAE_OK;
}

extern "C" int nvae_se_fwd(const float* t, const float* stat, const float* xres, int B, int HW, int C, int hid,
                           const float* w1, const float* b1, const float* w2, const float* b2, float alpha, float beta,
                           float* pooled, float* hidden, float* gate, float* y, nvae_stream_t stream) {
  int rc = se_check(B, HW, C, hid);
  if (rc) return rc;
  if (!t || !xres || !w1 || !b1 || !w2 || !b2 || !pooled || !hidden || !gate || !y) return NVAE_E_NULLPTR;
  const int S = se_cluster_size(B, HW, C);
  launch_cluster(se_fwd_cluster_kernel, dim3(B * S), kSeThreads, 0, stream, dim3(S, 1, 1), t, stat, HW, C, hid, w1, b1, w2, b2,
                 pooled, hidden, gate, xres, alpha, beta, y, S);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" size_t nvae_se_bwd_ws_bytes(int B, int C, int hid) { return (size_t)B * (2 * C + hid) * sizeof(float); }

extern "C" int nvae_se_bwd(const float* dy, const float* t, const float* stat, int B, int HW, int C, int hid,
                           const float* w1, const float* w2, const float* pooled, const float* hidden,
                           const float* gate, float alpha, float beta, float* dt, float* dxres, int dxres_accumulate,
                           float* dw1, float* db1, float* dw2, float* db2, void* ws, size_t ws_bytes,
                           nvae_stream_t stream) {
  int rc = se_check(B, HW, C, hid);
  if (rc) return rc;
  if (!dy || !t || !w1 || !w2 || !pooled || !hidden || !gate || !dt || !dw1 || !db1 || !dw2 || !db2)
    return NVAE_E_NULLPTR;
  if (ws == nullptr || ws_bytes < nvae_se_bwd_ws_bytes(B, C, hid)) return NVAE_E_WORKSPACE;
  float* dz2 = reinterpret_cast<float*>(ws);
  float* dpool = dz2 + (size_t)B * C;
  float* dh = dpool + (size_t)B * C;
  const int S = se_cluster_size(B, HW, C);
  launch_cluster(se_bwd_cluster_kernel, dim3(B * S), kSeThreads, 0, stream, dim3(S, 1, 1), dy, t, stat, HW, C, hid, w1, w2,
                 hidden, gate, beta, dz2, dh, dpool, alpha, dt, dxres, dxres_accumulate, S);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  const int total = 2 * C * hid + C + hid;
  nvae::launch(se_bwd_weights_kernel, (total * 8 + 127) / 128, 128, 0, stream, pooled, hidden, dz2, dh, B, C, hid, dw1, db1, dw2, db2);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}
