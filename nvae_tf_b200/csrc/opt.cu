// Optimizer (Adamax + CosineDecay, train.py:128-131), KL warm-up schedule (models.py:121-122),
// Philox epsilon generator (common.py:67) and small launch-only utilities.
#include "common.cuh"

namespace nvae {

// hyper layout: [0]=beta, [1]=lr_t (= lr/(1-b1^t)), [2]=lr, [3]=t (1-based), [4]=steps used for beta
__global__ void schedule_kernel(int64_t* counters, float* hyper, float warmup_iters, float lr0, float decay_steps,
                                float b1, int advance) {
  nvae::pdl_enter();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int64_t steps = counters[0], iters = counters[1];
  const float beta = warmup_iters > 0.f ? fminf((float)steps / warmup_iters, 1.f) : 1.f;
  const double tt = (double)(iters < (int64_t)decay_steps ? iters : (int64_t)decay_steps);
  const double lr = decay_steps > 0.f ? (double)lr0 * 0.5 * (1.0 + cos(3.14159265358979323846 * tt / (double)decay_steps))
                                      : (double)lr0;
  const double t = (double)(iters + 1);
  hyper[0] = beta;
  hyper[1] = (float)(lr / (1.0 - pow((double)b1, t)));
  hyper[2] = (float)lr;
  hyper[3] = (float)t;
  hyper[4] = (float)steps;
  if (advance & 1) counters[0] = steps + 1;
  if (advance & 2) counters[1] = iters + 1;
}

__global__ void adamax_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                              float* __restrict__ v, int64_t n4, const float* __restrict__ hyper, float b1, float b2,
                              float eps, float gs) {
  nvae::pdl_enter();
  const float lr_t = hyper[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pv = *reinterpret_cast<float4*>(p + i * 4), mv = *reinterpret_cast<float4*>(m + i * 4),
           vv = *reinterpret_cast<float4*>(v + i * 4);
    const float4 gv = ldg4(g + i * 4);
    const float gx = gv.x * gs, gy = gv.y * gs, gz = gv.z * gs, gw = gv.w * gs;
    mv.x = b1 * mv.x + (1.f - b1) * gx; mv.y = b1 * mv.y + (1.f - b1) * gy;
    mv.z = b1 * mv.z + (1.f - b1) * gz; mv.w = b1 * mv.w + (1.f - b1) * gw;
    vv.x = fmaxf(b2 * vv.x, fabsf(gx)); vv.y = fmaxf(b2 * vv.y, fabsf(gy));
    vv.z = fmaxf(b2 * vv.z, fabsf(gz)); vv.w = fmaxf(b2 * vv.w, fabsf(gw));
    pv.x -= lr_t * mv.x / (vv.x + eps); pv.y -= lr_t * mv.y / (vv.y + eps);
    pv.z -= lr_t * mv.z / (vv.z + eps); pv.w -= lr_t * mv.w / (vv.w + eps);
    stg4(p + i * 4, pv); stg4(m + i * 4, mv); stg4(v + i * 4, vv);
  }
}

__global__ void fill_kernel(float* p, int64_t n, float v) {
  nvae::pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void axpby_kernel(const float* __restrict__ x, float a, float* __restrict__ y, float b, int64_t n) {
  nvae::pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = b == 0.f ? a * x[i] : fmaf(a, x[i], b * y[i]);
}
__global__ void reparam_kernel(const float* __restrict__ mu, const float* __restrict__ sigma,
                               const float* __restrict__ eps, float sigma_scale, float* __restrict__ z, int64_t n) {
  nvae::pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    z[i] = fmaf(eps[i], sigma[i] * sigma_scale, mu[i]);
}
__global__ void broadcast_rows_kernel(const float* __restrict__ src, int64_t row, int B, float* __restrict__ dst) {
  nvae::pdl_enter();
  const int64_t n = row * B;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i % row];
}
__global__ void reduce_rows_kernel(const float* __restrict__ src, int64_t row, int B, float* __restrict__ dst) {
  nvae::pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < row; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += src[(int64_t)b * row + i];
    dst[i] = s;
  }
}

// Philox4x32-10 (Salmon et al. 2011), key = seed, counter = (index, step, stream_id)
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}
__global__ void philox_normal_kernel(float* __restrict__ out, int64_t n, uint64_t seed,
                                     const int64_t* __restrict__ counters, uint64_t stream_id) {
  nvae::pdl_enter();
  const uint64_t step = counters ? (uint64_t)counters[0] : 0ull;
  const int64_t n4 = (n + 3) / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)step,
                     (uint32_t)(stream_id ^ (step >> 32 << 16))};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c, k);
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; j += 2) {  // Box-Muller
      const float u1 = ((float)c[j] + 1.f) * 2.3283064365386963e-10f;  // (0,1]
      const float u2 = (float)c[j + 1] * 2.3283064365386963e-10f;
      const float rad = sqrtf(-2.f * logf(u1));
      float sn, cs;
      sincospif(2.f * u2, &sn, &cs);
      r[j] = rad * cs;
      r[j + 1] = rad * sn;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i * 4 + j < n) out[i * 4 + j] = r[j];
  }
}

// mode 0: sigmoid(l) (Bernoulli.probs_parameter / mean);  mode 1: U < sigmoid(l) (Bernoulli.sample)
__global__ void bernoulli_image_kernel(const float* __restrict__ logits, int64_t n, int mode, uint64_t seed,
                                       uint64_t stream_id, float* __restrict__ out) {
  nvae::pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float p = 1.f / (1.f + expf(-logits[i]));
    if (mode == 0) {
      out[i] = p;
    } else {
      uint32_t c[4] = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32), 0x5eedu, (uint32_t)stream_id};
      uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
      for (int r = 0; r < 10; ++r) philox_round(c, k);
      out[i] = (float)c[0] * 2.3283064365386963e-10f < p ? 1.f : 0.f;
    }
  }
}

static int grid_for(int64_t n, int threads) {
  int64_t b = ceil_div(n, threads);
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace nvae

using namespace nvae;

unsigned long long nvae_launch_counter = 0;
extern "C" uint64_t nvae_launch_count(void) { return (uint64_t)nvae_launch_counter; }
extern "C" int nvae_version(void) { return 100; }

extern "C" int nvae_graph_instantiate(void* graph, int use_node_priority, void** exec) {
  if (graph == nullptr || exec == nullptr) return NVAE_E_NULLPTR;
  cudaGraphExec_t e = nullptr;
  const cudaError_t rc = cudaGraphInstantiateWithFlags(&e, reinterpret_cast<cudaGraph_t>(graph),
                                                       use_node_priority ? cudaGraphInstantiateFlagUseNodePriority : 0);
  *exec = e;
  return (int)rc;
}
extern "C" int nvae_graph_launch(void* exec, nvae_stream_t stream) {
  if (exec == nullptr) return NVAE_E_NULLPTR;
  return (int)cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(exec), stream);
}
extern "C" int nvae_graph_destroy(void* exec) {
  return exec == nullptr ? NVAE_OK : (int)cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(exec));
}
extern "C" const char* nvae_build_info(void) { return "libnvae_b200 sm_100a (tcgen05/TMA) built " __DATE__ " " __TIME__; }

extern "C" int nvae_schedule_step(int64_t* counters, float* hyper, float warmup_iters, float lr0, float decay_steps,
                                  float b1, int advance, nvae_stream_t stream) {
  if (!counters || !hyper) return NVAE_E_NULLPTR;
  nvae::launch(schedule_kernel, 1, 32, 0, stream, counters, hyper, warmup_iters, lr0, decay_steps, b1, advance);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_adamax(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, float b1,
                           float b2, float eps, float grad_scale, nvae_stream_t stream) {
  if (n <= 0 || (n & 3)) return NVAE_E_BADSHAPE;
  if (!p || !g || !m || !v || !hyper) return NVAE_E_NULLPTR;
  nvae::launch(adamax_kernel, grid_for(n / 4, 256), 256, 0, stream, p, g, m, v, n / 4, hyper, b1, b2, eps, grad_scale);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_fill(float* p, int64_t n, float value, nvae_stream_t stream) {
  if (n <= 0) return NVAE_E_BADSHAPE;
  if (!p) return NVAE_E_NULLPTR;
  nvae::launch(fill_kernel, grid_for(n, 256), 256, 0, stream, p, n, value);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_axpby(const float* x, float a, float* y, float b, int64_t n, nvae_stream_t stream) {
  if (n <= 0) return NVAE_E_BADSHAPE;
  if (!x || !y) return NVAE_E_NULLPTR;
  nvae::launch(axpby_kernel, grid_for(n, 256), 256, 0, stream, x, a, y, b, n);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_reparam(const float* mu, const float* sigma, const float* eps, float sigma_scale, float* z,
                            int64_t n, nvae_stream_t stream) {
  if (n <= 0) return NVAE_E_BADSHAPE;
  if (!mu || !sigma || !eps || !z) return NVAE_E_NULLPTR;
  nvae::launch(reparam_kernel, grid_for(n, 256), 256, 0, stream, mu, sigma, eps, sigma_scale, z, n);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_broadcast_rows(const float* src, int64_t row_elems, int B, float* dst, nvae_stream_t stream) {
  if (row_elems <= 0 || B <= 0) return NVAE_E_BADSHAPE;
  if (!src || !dst) return NVAE_E_NULLPTR;
  nvae::launch(broadcast_rows_kernel, grid_for(row_elems * B, 256), 256, 0, stream, src, row_elems, B, dst);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_reduce_rows(const float* src, int64_t row_elems, int B, float* dst, nvae_stream_t stream) {
  if (row_elems <= 0 || B <= 0) return NVAE_E_BADSHAPE;
  if (!src || !dst) return NVAE_E_NULLPTR;
  nvae::launch(reduce_rows_kernel, grid_for(row_elems, 128), 128, 0, stream, src, row_elems, B, dst);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_philox_normal(float* out, int64_t n, uint64_t seed, const int64_t* counters, uint64_t stream_id,
                                  nvae_stream_t stream) {
  if (n <= 0) return NVAE_E_BADSHAPE;
  if (!out) return NVAE_E_NULLPTR;
  nvae::launch(philox_normal_kernel, grid_for((n + 3) / 4, 256), 256, 0, stream, out, n, seed, counters, stream_id);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_bernoulli_image(const float* logits, int64_t n, int mode, uint64_t seed, uint64_t stream_id,
                                    float* out, nvae_stream_t stream) {
  if (n <= 0 || (mode != 0 && mode != 1)) return NVAE_E_BADSHAPE;
  if (!logits || !out) return NVAE_E_NULLPTR;
  nvae::launch(bernoulli_image_kernel, grid_for(n, 256), 256, 0, stream, logits, n, mode, seed, stream_id, out);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

extern "C" int nvae_l2_flush(float* scratch, int64_t n, nvae_stream_t stream) {
  return nvae_fill(scratch, n, 0.f, stream);
}
