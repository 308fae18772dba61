// tfa.layers.SpectralNormalization(power_iterations=1) for ALL wrapped Conv2D layers in four
// launches (the reference issues ~8 TF ops per layer x 163 layers), plus packing of the normalised
// kernels into the TF32 operand layouts of the tcgen05 convolutions.  Semantics:
//   v = l2n(u W^T); u' = l2n(v W); sigma = (v W) u'^T; W <- W / sigma; u <- u'
// HBM-bound multi-tensor kernels over a flat parameter arena: work is cut into 64-row chunks of
// [rows, cout] matrices so 148 SMs stay busy across layers from 9x32 to 9600x384.
#include "common.cuh"

namespace nvae {

constexpr int kSnThreads = 256;
constexpr int kSnRows = NVAE_SN_ROWS_PER_CHUNK;

// K1: v_raw[row] = sum_co W[row,co]*u[co]
__global__ void __launch_bounds__(kSnThreads) sn_wu_kernel(const float* __restrict__ params,
                                                           const float* __restrict__ state,
                                                           const NvaeSnLayer* __restrict__ layers,
                                                           const int32_t* __restrict__ chunk_layer,
                                                           float* __restrict__ ws) {
  nvae::pdl_enter();
  const NvaeSnLayer L = layers[chunk_layer[blockIdx.x]];
  const int chunk = blockIdx.x - L.chunk0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* W = params + L.w_off;
  const float* u = state + L.u_off;
  const int r0 = chunk * kSnRows;
  for (int r = r0 + warp; r < r0 + kSnRows && r < L.rows; r += kSnThreads / 32) {
    const float* wr = W + (int64_t)r * L.cout;
    float s = 0.f;
    for (int c = lane; c < L.cout; c += 32) s = fmaf(wr[c], u[c], s);
    s = warp_sum(s);
    if (lane == 0) ws[L.v_off + r] = s;
  }
}

// K2: t_partial[chunk][co] = sum_{r in chunk} v_raw[r]*W[r,co];  slot [cout] of the partial = sum v_raw^2
__global__ void __launch_bounds__(kSnThreads) sn_vw_kernel(const float* __restrict__ params,
                                                           const NvaeSnLayer* __restrict__ layers,
                                                           const int32_t* __restrict__ chunk_layer,
                                                           float* __restrict__ ws) {
  nvae::pdl_enter();
  __shared__ float sv[kSnRows];
  const NvaeSnLayer L = layers[chunk_layer[blockIdx.x]];
  const int chunk = blockIdx.x - L.chunk0;
  const float* W = params + L.w_off;
  const int r0 = chunk * kSnRows;
  const int nr = (L.rows - r0) < kSnRows ? (L.rows - r0) : kSnRows;
  if (threadIdx.x < kSnRows) sv[threadIdx.x] = threadIdx.x < nr ? ws[L.v_off + r0 + threadIdx.x] : 0.f;
  __syncthreads();
  float* tp = ws + L.t_off + (int64_t)chunk * (L.cout + 1);
  for (int c = threadIdx.x; c < L.cout; c += kSnThreads) {
    float s = 0.f;
    for (int r = 0; r < nr; ++r) s = fmaf(sv[r], W[(int64_t)(r0 + r) * L.cout + c], s);
    tp[c] = s;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int r = 0; r < nr; ++r) s = fmaf(sv[r], sv[r], s);
    tp[L.cout] = s;
  }
}

// K3 (one CTA per layer): combine partials in fixed order, u <- u', sigma
__global__ void __launch_bounds__(kSnThreads) sn_finalize_kernel(float* __restrict__ state,
                                                                 const NvaeSnLayer* __restrict__ layers,
                                                                 float* __restrict__ ws, float* __restrict__ sigma) {
  nvae::pdl_enter();
  __shared__ float red[33];
  const NvaeSnLayer L = layers[blockIdx.x];
  float nv2 = 0.f;
  for (int k = 0; k < L.n_chunks; ++k) nv2 += ws[L.t_off + (int64_t)k * (L.cout + 1) + L.cout];
  const float inv_nv = rsqrtf(fmaxf(nv2, 1e-12f));
  float* t0 = ws + L.t_off;  // reuse chunk-0 partial row as storage for t
  float nt2 = 0.f;
  for (int c = threadIdx.x; c < L.cout; c += kSnThreads) {
    float s = 0.f;
    for (int k = 0; k < L.n_chunks; ++k) s += ws[L.t_off + (int64_t)k * (L.cout + 1) + c];
    s *= inv_nv;  // t = v W with v = v_raw / |v_raw|
    t0[c] = s;
    nt2 = fmaf(s, s, nt2);
  }
  nt2 = block_sum(nt2, red);
  const float inv_nt = rsqrtf(fmaxf(nt2, 1e-12f));
  float sg = 0.f;
  for (int c = threadIdx.x; c < L.cout; c += kSnThreads) {
    const float t = t0[c], un = t * inv_nt;
    state[L.u_off + c] = un;
    sg = fmaf(t, un, sg);
  }
  sg = block_sum(sg, red);
  if (threadIdx.x == 0) sigma[blockIdx.x] = sg;
}

// K4: W <- W/sigma (in place), TF32-rounded HWIO copy (dgrad B operand), transposed TF32 copy
// [cout_pad][taps][cin_pad] (fwd B operand, K-major).  pack_exact: the copies keep the full fp32 values (3xTF32
// mode splits them into high and low parts inside the convolution kernel).  Transposition goes through shared memory
// so both the read (co fastest) and the write (ci fastest) are coalesced.
__global__ void __launch_bounds__(kSnThreads) sn_scale_pack_kernel(float* __restrict__ params, float* __restrict__ pack,
                                                                   const NvaeSnLayer* __restrict__ layers,
                                                                   const int32_t* __restrict__ chunk_layer,
                                                                   const float* __restrict__ sigma, int power_iter,
                                                                   int pack_exact) {
  nvae::pdl_enter();
  constexpr int kTC = 64;  // columns per transposition tile
  __shared__ float tile[kSnRows][kTC + 1];
  const int li = chunk_layer[blockIdx.x];
  const NvaeSnLayer L = layers[li];
  const int chunk = blockIdx.x - L.chunk0;
  float* W = params + L.w_off;
  const float inv_sigma = power_iter ? 1.f / sigma[li] : 1.f;
  const int r0 = chunk * kSnRows;
  const int nr = (L.rows - r0) < kSnRows ? (L.rows - r0) : kSnRows;
  const int K = L.taps * L.cin_pad;
  // Phase 1: the chunk's rows are one contiguous block of nr*cout floats -> streaming in-place scale (and the
  // HWIO operand copy), 128-bit accesses, four independent loads in flight per thread.
  if (power_iter || L.rnd_off >= 0) {
    const int64_t base = (int64_t)r0 * L.cout, n = (int64_t)nr * L.cout;
    float* wb = W + base;
    float* pb = L.rnd_off >= 0 ? pack + L.rnd_off + base : nullptr;
    if ((((uintptr_t)wb | (uintptr_t)pb) & 15u) == 0 && (n & 3) == 0) {
      const int64_t n4 = n >> 2;
      for (int64_t i0 = threadIdx.x; i0 < n4; i0 += 4 * kSnThreads) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t i = i0 + (int64_t)j * kSnThreads;
          v[j] = i < n4 ? *reinterpret_cast<const float4*>(wb + 4 * i) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t i = i0 + (int64_t)j * kSnThreads;
          if (i >= n4) continue;
          v[j].x *= inv_sigma; v[j].y *= inv_sigma; v[j].z *= inv_sigma; v[j].w *= inv_sigma;
          if (power_iter) stg4(wb + 4 * i, v[j]);
          if (pb != nullptr)
            stg4(pb + 4 * i, pack_exact ? v[j] : make_float4(round_tf32(v[j].x), round_tf32(v[j].y), round_tf32(v[j].z),
                                                                 round_tf32(v[j].w)));
        }
      }
    } else {
      for (int64_t i = threadIdx.x; i < n; i += kSnThreads) {
        const float v = wb[i] * inv_sigma;
        if (power_iter) wb[i] = v;
        if (pb != nullptr) pb[i] = pack_exact ? v : round_tf32(v);
      }
    }
    __syncthreads();  // phase 2 re-reads the scaled block (same CTA, global memory)
  }
  if (L.tr_off < 0) return;
  // Phase 2: transposed operand copy [cout][taps][cin_pad] through shared memory (reads co-fastest, writes ci-fastest)
  for (int cb = 0; cb < L.cout; cb += kTC) {
    for (int i = threadIdx.x; i < kSnRows * kTC; i += kSnThreads) {
      const int r = i / kTC, cc = i - r * kTC, c = cb + cc;
      tile[r][cc] = (r < nr && c < L.cout) ? W[(int64_t)(r0 + r) * L.cout + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSnRows * kTC; i += kSnThreads) {
      const int r = i % kSnRows, cc = i / kSnRows, c = cb + cc;
      if (r < nr && c < L.cout) {
        const int row = r0 + r, tap = row / L.cin, ci = row - tap * L.cin;
        const float v = tile[r][cc];
        pack[L.tr_off + (int64_t)c * K + (int64_t)tap * L.cin_pad + ci] = pack_exact ? v : round_tf32(v);
      }
    }
    __syncthreads();
  }
}

}  // namespace nvae

using namespace nvae;

extern "C" int nvae_spectral_norm(float* params, float* state, float* pack, const NvaeSnLayer* layers_dev,
                                  int n_layers, const int32_t* chunk_layer_dev, int n_chunks_total, int power_iter,
                                  int pack_exact, float* sigma_out, float* ws, nvae_stream_t stream) {
  if (n_layers <= 0 || n_chunks_total <= 0) return NVAE_E_BADSHAPE;
  if (!params || !layers_dev || !chunk_layer_dev) return NVAE_E_NULLPTR;
  if (pack_exact && pack == nullptr) return NVAE_E_NULLPTR;
  if (power_iter) {
    if (!state || !sigma_out || !ws) return NVAE_E_NULLPTR;
    nvae::launch(sn_wu_kernel, n_chunks_total, kSnThreads, 0, stream, params, state, layers_dev, chunk_layer_dev, ws);
    NVAE_RETURN_IF_LAUNCH_FAILED();
    nvae::launch(sn_vw_kernel, n_chunks_total, kSnThreads, 0, stream, params, layers_dev, chunk_layer_dev, ws);
    NVAE_RETURN_IF_LAUNCH_FAILED();
    nvae::launch(sn_finalize_kernel, n_layers, kSnThreads, 0, stream, state, layers_dev, ws, sigma_out);
    NVAE_RETURN_IF_LAUNCH_FAILED();
  }
  if (power_iter || pack != nullptr) {
    nvae::launch(sn_scale_pack_kernel, n_chunks_total, kSnThreads, 0, stream, params, pack, layers_dev, chunk_layer_dev,
                                                                   sigma_out, power_iter, pack_exact);
    NVAE_RETURN_IF_LAUNCH_FAILED();
  }
  return NVAE_OK;
}
