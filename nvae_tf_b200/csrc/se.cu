// Squeeze-and-Excitation gate + residual merge (common.py:129-142 fused with encoder.py:107,
// decoder.py:147, preprocess.py:107, postprocess.py:58).  Bandwidth-bound: one pass over t for the
// global pool (+ the two tiny dense layers in the same CTA), one pass for the gated residual.
#include "common.cuh"

namespace nvae {

constexpr int kSeThreads = 256;
constexpr int kSeMaxC = 1024;
constexpr int kSeMaxHid = 64;
constexpr size_t kSeFuseBytes = 64 * 1024;  // per-sample tensor size up to which the per-sample CTA also applies the gate

// One CTA per sample: pooled' = affine(mean_hw t), hidden = relu(W1^T pooled' + b1),
// gate = sigmoid(W2^T hidden + b2).
__global__ void __launch_bounds__(kSeThreads) se_pool_gate_kernel(
    const float* __restrict__ t, const float* __restrict__ stat, int HW, int C, int hid, const float* __restrict__ w1,
    const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
    float* __restrict__ pooled, float* __restrict__ hidden, float* __restrict__ gate, const float* __restrict__ xres,
    float alpha, float beta, float* __restrict__ y) {
  nvae::pdl_enter();
  __shared__ float part[kSeThreads * 4];
  __shared__ __align__(16) float spool[kSeMaxC];
  __shared__ float shid[kSeMaxHid];
  const int b = blockIdx.x, C4 = C >> 2, tid = threadIdx.x;
  const int G = kSeThreads / C4;  // row groups
  const int grp = tid / C4, c4 = tid % C4;
  const float* tb = t + (int64_t)b * HW * C;
  float4 acc = make_float4(0, 0, 0, 0);
  if (grp < G) {
    int r = grp;
    for (; r + 3 * G < HW; r += 4 * G) {
      const float4 v0 = ldg4(tb + (int64_t)r * C + c4 * 4), v1 = ldg4(tb + (int64_t)(r + G) * C + c4 * 4),
                   v2 = ldg4(tb + (int64_t)(r + 2 * G) * C + c4 * 4), v3 = ldg4(tb + (int64_t)(r + 3 * G) * C + c4 * 4);
      acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
      acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; r < HW; r += G) {
      const float4 v = ldg4(tb + (int64_t)r * C + c4 * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(&part[tid * 4]) = acc;
  }
  __syncthreads();
  for (int c = tid; c < C; c += kSeThreads) {
    float s = 0.f;
    for (int g = 0; g < G; ++g) s += part[(g * C4 + (c >> 2)) * 4 + (c & 3)];
    s *= 1.f / (float)HW;
    if (stat != nullptr) s = fmaf(s, stat[2 * C + c], stat[3 * C + c]);
    spool[c] = s;
    pooled[(int64_t)b * C + c] = s;
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int j = warp; j < hid; j += kSeThreads / 32) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(spool[c], __ldg(w1 + (int64_t)c * hid + j), s);
    s = warp_sum(s);
    if (lane == 0) {
      s = fmaxf(s + b1[j], 0.f);
      shid[j] = s;
      hidden[(int64_t)b * hid + j] = s;
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += kSeThreads) {
    float s = b2[c];
    for (int j = 0; j < hid; ++j) s = fmaf(shid[j], __ldg(w2 + (int64_t)j * C + c), s);
    s = 1.f / (1.f + expf(-s));
    gate[(int64_t)b * C + c] = s;
    spool[c] = s;  // (the pooled values are no longer needed)
  }
  if (y == nullptr) return;
  // fused residual merge for small samples: y = alpha*xres + beta*t'*gate -- the sample's t was just streamed by this
  // CTA (L1/L2 hits), and one launch disappears
  __syncthreads();
  const float* xb = xres + (int64_t)b * HW * C;
  float* yb = y + (int64_t)b * HW * C;
  const int n4 = HW * C4;
  for (int i = tid; i < n4; i += kSeThreads) {
    const int k4 = i % C4;
    float4 v = ldg4(tb + (int64_t)i * 4);
    if (stat != nullptr) {
      const float4 sc = ldg4(stat + 2 * C + k4 * 4), sh = ldg4(stat + 3 * C + k4 * 4);
      v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    }
    const float4 g = *reinterpret_cast<const float4*>(&spool[k4 * 4]);
    const float4 xr = ldg4(xb + (int64_t)i * 4);
    float4 o;
    o.x = fmaf(alpha, xr.x, beta * v.x * g.x); o.y = fmaf(alpha, xr.y, beta * v.y * g.y);
    o.z = fmaf(alpha, xr.z, beta * v.z * g.z); o.w = fmaf(alpha, xr.w, beta * v.w * g.w);
    stg4(yb + (int64_t)i * 4, o);
  }
}

// y = alpha*xres + beta*t'*gate[b,c]
__global__ void se_apply_kernel(const float* __restrict__ t, const float* __restrict__ stat,
                                const float* __restrict__ xres, const float* __restrict__ gate, int64_t n4, int C4,
                                int64_t hwc4, float alpha, float beta, float* __restrict__ y) {
  nvae::pdl_enter();
  const int C = C4 * 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const int64_t b = i / hwc4;
    float4 v = ldg4(t + i * 4);
    if (stat != nullptr) {
      const float4 sc = ldg4(stat + 2 * C + c4 * 4), sh = ldg4(stat + 3 * C + c4 * 4);
      v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    }
    const float4 g = ldg4(gate + b * C + c4 * 4);
    const float4 xr = ldg4(xres + i * 4);
    float4 o;
    o.x = fmaf(alpha, xr.x, beta * v.x * g.x); o.y = fmaf(alpha, xr.y, beta * v.y * g.y);
    o.z = fmaf(alpha, xr.z, beta * v.z * g.z); o.w = fmaf(alpha, xr.w, beta * v.w * g.w);
    stg4(y + i * 4, o);
  }
}

// One CTA per sample: r[c] = beta*sum_hw dy*t'; chain through sigmoid / dense2 / relu / dense1.
// Writes dz2 [B,C], dh [B,hid], dpool [B,C].
__global__ void __launch_bounds__(kSeThreads) se_bwd_gate_kernel(
    const float* __restrict__ dy, const float* __restrict__ t, const float* __restrict__ stat, int HW, int C, int hid,
    const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ hidden,
    const float* __restrict__ gate, float beta, float* __restrict__ dz2, float* __restrict__ dh,
    float* __restrict__ dpool, float alpha, float* __restrict__ dt, float* __restrict__ dxres, int dxres_accumulate) {
  nvae::pdl_enter();
  __shared__ float part[kSeThreads * 4];
  __shared__ __align__(16) float sdz[kSeMaxC];
  __shared__ float sdh[kSeMaxHid];
  const int b = blockIdx.x, C4 = C >> 2, tid = threadIdx.x;
  const int G = kSeThreads / C4;
  const int grp = tid / C4, c4 = tid % C4;
  const float* tb = t + (int64_t)b * HW * C;
  const float* db = dy + (int64_t)b * HW * C;
  float4 acc = make_float4(0, 0, 0, 0);
  if (grp < G) {
    float4 sc = make_float4(1, 1, 1, 1), sh = make_float4(0, 0, 0, 0);
    if (stat != nullptr) {
      sc = ldg4(stat + 2 * C + c4 * 4);
      sh = ldg4(stat + 3 * C + c4 * 4);
    }
#pragma unroll 4
    for (int r = grp; r < HW; r += G) {
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
    const float gt = gate[(int64_t)b * C + c];
    const float d = beta * s * gt * (1.f - gt);
    sdz[c] = d;
    dz2[(int64_t)b * C + c] = d;
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int j = warp; j < hid; j += kSeThreads / 32) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(sdz[c], __ldg(w2 + (int64_t)j * C + c), s);
    s = warp_sum(s);
    if (lane == 0) {
      s = hidden[(int64_t)b * hid + j] > 0.f ? s : 0.f;
      sdh[j] = s;
      dh[(int64_t)b * hid + j] = s;
    }
  }
  __syncthreads();
  __syncthreads();  // (sdz is reused below: everyone is done reading it)
  for (int c = tid; c < C; c += kSeThreads) {
    float s = 0.f;
    for (int j = 0; j < hid; ++j) s = fmaf(sdh[j], __ldg(w1 + (int64_t)c * hid + j), s);
    dpool[(int64_t)b * C + c] = s;
    sdz[c] = s;
  }
  if (dt == nullptr) return;
  // fused apply for small samples: dt' = beta*dy*gate + dpool/HW ; dxres (+)= alpha*dy
  __syncthreads();
  const float inv_hw = 1.f / (float)HW;
  const float* gb = gate + (int64_t)b * C;
  float* dtb = dt + (int64_t)b * HW * C;
  float* dxb = dxres != nullptr ? dxres + (int64_t)b * HW * C : nullptr;
  const int n4 = HW * C4;
  for (int i = tid; i < n4; i += kSeThreads) {
    const int k4 = i % C4;
    const float4 d = ldg4(db + (int64_t)i * 4), g = ldg4(gb + k4 * 4);
    const float4 p = *reinterpret_cast<const float4*>(&sdz[k4 * 4]);
    float4 o;
    o.x = fmaf(beta * d.x, g.x, p.x * inv_hw); o.y = fmaf(beta * d.y, g.y, p.y * inv_hw);
    o.z = fmaf(beta * d.z, g.z, p.z * inv_hw); o.w = fmaf(beta * d.w, g.w, p.w * inv_hw);
    stg4(dtb + (int64_t)i * 4, o);
    if (dxb != nullptr) {
      float4 r = make_float4(alpha * d.x, alpha * d.y, alpha * d.z, alpha * d.w);
      if (dxres_accumulate) {
        const float4 e = *reinterpret_cast<const float4*>(dxb + (int64_t)i * 4);
        r.x += e.x; r.y += e.y; r.z += e.z; r.w += e.w;
      }
      stg4(dxb + (int64_t)i * 4, r);
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

// dt' = beta*dy*gate + dpool/HW ; dxres (+)= alpha*dy
__global__ void se_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ gate,
                                    const float* __restrict__ dpool, int64_t n4, int C4, int64_t hwc4, float inv_hw,
                                    float alpha, float beta, float* __restrict__ dt, float* __restrict__ dxres,
                                    int dxres_accumulate) {
  nvae::pdl_enter();
  const int C = C4 * 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const int64_t b = i / hwc4;
    const float4 d = ldg4(dy + i * 4), g = ldg4(gate + b * C + c4 * 4), p = ldg4(dpool + b * C + c4 * 4);
    float4 o;
    o.x = fmaf(beta * d.x, g.x, p.x * inv_hw); o.y = fmaf(beta * d.y, g.y, p.y * inv_hw);
    o.z = fmaf(beta * d.z, g.z, p.z * inv_hw); o.w = fmaf(beta * d.w, g.w, p.w * inv_hw);
    stg4(dt + i * 4, o);
    if (dxres != nullptr) {
      float4 r = make_float4(alpha * d.x, alpha * d.y, alpha * d.z, alpha * d.w);
      if (dxres_accumulate) {
        const float4 e = *reinterpret_cast<const float4*>(dxres + i * 4);
        r.x += e.x; r.y += e.y; r.z += e.z; r.w += e.w;
      }
      stg4(dxres + i * 4, r);
    }
  }
}

static int se_grid(int64_t n, int threads) {
  int64_t b = ceil_div(n, threads);
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace nvae

using namespace nvae;

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
  // one CTA per sample also merges the residual when a sample is small (<= 64 KB): at the model's 4x4 / 8x8 / 16x16
  // scales that is one launch instead of two and the second read of t never leaves the SM's cache
  const bool fuse = (size_t)HW * C * sizeof(float) <= kSeFuseBytes;
  nvae::launch(se_pool_gate_kernel, B, kSeThreads, 0, stream, t, stat, HW, C, hid, w1, b1, w2, b2, pooled, hidden, gate, xres,
               alpha, beta, fuse ? y : (float*)nullptr);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  if (fuse) return NVAE_OK;
  const int64_t n4 = (int64_t)B * HW * (C / 4);
  nvae::launch(se_apply_kernel, se_grid(n4, 256), 256, 0, stream, t, stat, xres, gate, n4, C / 4, (int64_t)HW * (C / 4), alpha,
                                                        beta, y);
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
  const bool fuse = (size_t)HW * C * sizeof(float) <= kSeFuseBytes;
  nvae::launch(se_bwd_gate_kernel, B, kSeThreads, 0, stream, dy, t, stat, HW, C, hid, w1, w2, hidden, gate, beta, dz2, dh, dpool,
               alpha, fuse ? dt : (float*)nullptr, dxres, dxres_accumulate);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  const int total = 2 * C * hid + C + hid;
  nvae::launch(se_bwd_weights_kernel, (total * 8 + 127) / 128, 128, 0, stream, pooled, hidden, dz2, dh, B, C, hid, dw1, db1, dw2, db2);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  if (fuse) return NVAE_OK;
  const int64_t n4 = (int64_t)B * HW * (C / 4);
  nvae::launch(se_bwd_apply_kernel, se_grid(n4, 256), 256, 0, stream, dy, gate, dpool, n4, C / 4, (int64_t)HW * (C / 4),
                                                            1.f / (float)HW, alpha, beta, dt, dxres, dxres_accumulate);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}
