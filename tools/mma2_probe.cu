// Development probe (not product code): does tcgen05.mma.cta_group::2 with A read from TMEM and the B tile split over
// the two CTAs' shared memories compute what we think it does?  One cluster of 2 CTAs:
//   D[256 x N] = A[256 x 32] * B[N x 32]^T   (kind::tf32, small integers -> exact)
// CTA r holds rows [128r, 128r+128) of A in ITS tensor memory (tcgen05.st, lanes = rows, 32 K columns), rows
// [r*N/2, (r+1)*N/2) of B in ITS shared memory (K-major, 128-byte swizzle), and receives rows [128r, +128) of D in its
// tensor memory.  The leader (rank 0) issues the MMAs and commits to a barrier in both CTAs.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma2_probe tools/mma2_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(16u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int use_2cta) {
  __shared__ __align__(1024) uint8_t bsm[16384];  // up to 128 rows x 128 B
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_base;
  const uint32_t rank = cluster_rank();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // B half: rows [rank*N/2, +N/2), element (n,k) at n*128 + ((k/4) ^ (n&7))*16 + (k%4)*4
  for (int i = threadIdx.x; i < (N / 2) * 32; i += blockDim.x) {
    const int n = i / 32, k = i % 32;
    *reinterpret_cast<float*>(bsm + n * 128 + (((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4) = B[(rank * (N / 2) + n) * 32 + k];
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  // A rows -> TMEM lanes, columns [256, 288)
  {
    const int row = warp * 32 + lane;
    uint32_t v[32];
    for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(A[(rank * 128 + row) * 32 + k]);
    const uint32_t ta = tmem + 256u + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
        "%25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(ta), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
          "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
          "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();  // both CTAs: barrier initialised, B in smem, A in TMEM
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (rank == 0 && threadIdx.x == 0) {
    const uint64_t db = desc_sw128(smem_u32(bsm));
    const uint32_t idesc = idesc_tf32(256, N);
    for (int j = 0; j < 4; ++j) {
      const uint32_t acc = j > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
          ::"r"(tmem), "r"(tmem + 256u + 8u * j), "l"(db + (uint64_t)(2 * j)), "r"(idesc), "r"(acc)
          : "memory");
    }
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&done)), "h"(mask) : "memory");
  }
  while (!mbar_try_wait(smem_u32(&done), 0)) {}
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int row = warp * 32 + lane;
    for (int j = 0; j < N; j += 32) {
      uint32_t r[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)j)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int q = 0; q < 32; ++q) D[(rank * 128 + row) * N + j + q] = __uint_as_float(r[q]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  constexpr int N = 128;
  float *hA = (float*)malloc(256 * 32 * 4), *hB = (float*)malloc(N * 32 * 4), *hD = (float*)malloc(256 * N * 4);
  for (int i = 0; i < 256 * 32; ++i) hA[i] = (float)((i * 7 + i / 32) % 11 - 5);
  for (int i = 0; i < N * 32; ++i) hB[i] = (float)((i * 5 + i / 32 * 3) % 7 - 3);
  float *A, *B, *D;
  cudaMalloc(&A, 256 * 32 * 4); cudaMalloc(&B, N * 32 * 4); cudaMalloc(&D, 256 * N * 4);
  cudaMemcpy(A, hA, 256 * 32 * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(B, hB, N * 32 * 4, cudaMemcpyHostToDevice);
  cudaMemset(D, 0xFF, 256 * N * 4);
  probe<N><<<2, 128>>>(A, B, D, 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  cudaMemcpy(hD, D, 256 * N * 4, cudaMemcpyDeviceToHost);
  int bad = 0, first = -1;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0;
      for (int k = 0; k < 32; ++k) ref += hA[m * 32 + k] * hB[n * 32 + k];
      if (hD[m * N + n] != ref) { if (first < 0) first = m * N + n; ++bad; }
    }
  printf("mismatches: %d of %d", bad, 256 * N);
  if (first >= 0) {
    printf("  first at m=%d n=%d: got %g", first / N, first % N, hD[first]);
    // which (m', n') would explain it?
    for (int m = 0; m < 256 && first >= 0; ++m)
      for (int n = 0; n < N; ++n) {
        float ref = 0;
        for (int k = 0; k < 32; ++k) ref += hA[m * 32 + k] * hB[n * 32 + k];
        if (ref == hD[first] && ref != 0) { printf(" (== D_ref[%d][%d])", m, n); m = 256; break; }
      }
  }
  printf("\n");
  // per-quadrant mismatch map (rows 0-127/128-255 x cols 0-63/64-127)
  for (int qm = 0; qm < 2; ++qm)
    for (int qn = 0; qn < 2; ++qn) {
      int b = 0;
      for (int m = qm * 128; m < qm * 128 + 128; ++m)
        for (int n = qn * (N / 2); n < (qn + 1) * (N / 2); ++n) {
          float ref = 0;
          for (int k = 0; k < 32; ++k) ref += hA[m * 32 + k] * hB[n * 32 + k];
          b += hD[m * N + n] != ref;
        }
      printf("  rows %3d.. cols %3d..: %d bad\n", qm * 128, qn * (N / 2), b);
    }
  return 0;
}
