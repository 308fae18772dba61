// Shared device helpers for libnvae_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nvae_b200.h"

// every kernel launch in the library is followed by this macro: it also counts the launch (nvae_launch_count)
extern unsigned long long nvae_launch_counter;
#define NVAE_RETURN_IF_LAUNCH_FAILED()                 \
  do {                                                 \
    ++nvae_launch_counter;                             \
    cudaError_t e__ = cudaGetLastError();              \
    if (e__ != cudaSuccess) return (int)e__;           \
  } while (0)

#define NVAE_CUDA_TRY(expr)                            \
  do {                                                 \
    cudaError_t e__ = (expr);                          \
    if (e__ != cudaSuccess) return (int)e__;           \
  } while (0)

namespace nvae {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---- programmatic dependent launch (opt-in: NVAE_PDL=1) -------------------------------------------
// Every kernel begins with pdl_enter(): `launch_dependents` lets the NEXT kernel of the stream be scheduled (its
// CTAs become resident and run their prologue) as soon as every CTA of this one has started, `wait` blocks until the
// PREVIOUS kernel has completed and flushed its memory.  Because every kernel waits before its first dependent
// access and a kernel only completes after its own wait returned, stream order is preserved transitively.  Without
// the launch attribute both instructions are no-ops, which is the default: measured inside the step's CUDA graph the
// attribute gained nothing (39.9 ms without, 40.4 ms with -- graph kernel->kernel edges are already that cheap).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_trigger();
  pdl_wait();
}

bool pdl_enabled();  // conv.cu: reads NVAE_PDL once

// `cluster` = thread-block-cluster dimensions (grid must be a multiple), {1,1,1} = none
template <typename... KArgs, typename... Args>
inline void launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                           dim3 cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster.x * cluster.y * cluster.z > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster.x;
    attr[n].val.clusterDim.y = cluster.y;
    attr[n].val.clusterDim.z = cluster.z;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  (void)cudaLaunchKernelEx(&cfg, kernel, KArgs(static_cast<Args&&>(args))...);  // error picked up by cudaGetLastError
}
template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  launch_cluster(kernel, grid, block, smem, stream, dim3(1, 1, 1), static_cast<Args&&>(args)...);
}

// ---- thread-block clusters: barrier + distributed-shared-memory loads ---------------------------
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double dsmem_ld_f64(const double* local, unsigned rank) {
  double v;
  asm volatile(
      "{\n\t.reg .u32 la, ra;\n\t"
      "cvt.u32.u64 la, %1;\n\t"
      "mapa.shared::cluster.u32 ra, la, %2;\n\t"
      "ld.shared::cluster.f64 %0, [ra];\n\t}"
      : "=d"(v)
      : "l"(__cvta_generic_to_shared(local)), "r"(rank)
      : "memory");
  return v;
}

__device__ __forceinline__ float dsmem_ld_f32(const float* local, unsigned rank) {
  float v;
  asm volatile(
      "{\n\t.reg .u32 la, ra;\n\t"
      "cvt.u32.u64 la, %1;\n\t"
      "mapa.shared::cluster.u32 ra, la, %2;\n\t"
      "ld.shared::cluster.f32 %0, [ra];\n\t}"
      : "=f"(v)
      : "l"(__cvta_generic_to_shared(local)), "r"(rank)
      : "memory");
  return v;
}

// ---- activations -------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ float act_fwd(float u) {
  // ex2.approx + rcp.rn: ~1e-6 relative on the swish value, a quarter of the instructions of expf + IEEE divide
  // (the bandwidth-bound kernels that apply it are otherwise instruction-issue-bound)
  if (ACT == NVAE_ACT_SWISH) return u * __frcp_rn(1.f + __expf(-u));
  if (ACT == NVAE_ACT_ELU) return u > 0.f ? u : expm1f(u);
  return u;
}
template <int ACT>
__device__ __forceinline__ float act_grad(float u) {
  if (ACT == NVAE_ACT_SWISH) {
    const float s = __frcp_rn(1.f + __expf(-u));
    return s * (1.f + u * (1.f - s));
  }
  if (ACT == NVAE_ACT_ELU) return u > 0.f ? 1.f : expf(u);
  return 1.f;
}
__device__ __forceinline__ float act_fwd_rt(float u, int act) {
  return act == NVAE_ACT_SWISH ? act_fwd<NVAE_ACT_SWISH>(u) : act == NVAE_ACT_ELU ? act_fwd<NVAE_ACT_ELU>(u) : u;
}
__device__ __forceinline__ float act_grad_rt(float u, int act) {
  return act == NVAE_ACT_SWISH ? act_grad<NVAE_ACT_SWISH>(u) : act == NVAE_ACT_ELU ? act_grad<NVAE_ACT_ELU>(u) : 1.f;
}

// round-to-nearest TF32 (10-bit mantissa); low 13 bits of the result are zero
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ---- warp / block reductions ----------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Deterministic block sum (fixed tree); result valid in every thread. `red` needs 33 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// ---- 128-bit streaming access -----------------------------------------------------------------
// "Last block finalizes": after a CTA has written its partial results, one thread takes a ticket; the CTA that draws
// the last ticket of its group knows every partial is in L2 (fence before the ticket, fence after) and does the
// combine -- in a FIXED order over the partials, so the result does not depend on which CTA came last.  No CTA ever waits
// for another one (no co-residency assumption, no deadlock).  `ticket` must be zero at launch (a memset node ahead of it).
__device__ __forceinline__ bool last_block_of(unsigned* ticket, unsigned group_size) {
  __shared__ unsigned s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == group_size - 1u;
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stg4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace nvae
