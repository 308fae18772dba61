// Development probe (not product code): tcgen05.mma kind::tf32 issue throughput from shared-memory operands
// as a function of N and operand major-ness.  One CTA per SM, operands are whatever the shared memory holds
// (zero-initialised), NITER back-to-back MMAs into one TMEM accumulator, timed with clock64.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu
//   run  : tools/mma_probe
#include <cstdint>
#include <cstdio>
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
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint64_t layout) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ uint32_t idesc_tf32(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ uint32_t idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// mode 0: tf32 K-major SW128; 1: tf32 MN-major; 2: bf16 K-major SW128; 3: tf32 K-major SW32; 4: tf32 K-major SW64
// hammer: warps 1..3 stream float4 read-modify-writes over a separate 32 KB region while the MMAs run
__global__ void __launch_bounds__(128, 1) probe(int N, int mode, int niter, int hammer, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  __shared__ volatile int done;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 196 * 1024 / 4; i += blockDim.x) {
    // hammer==2: random finite floats in [1,2) (realistic toggling / power); else zeros
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    reinterpret_cast<uint32_t*>(smem_raw)[i] = hammer == 2 ? (0x3F800000u | (h & 0x007FFFFFu)) : 0u;
  }
  if (threadIdx.x == 0) {
    done = 0;
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    const uint32_t sa = base, sb = base + 16384;
    uint64_t da, db;
    uint32_t idesc;
    if (mode == 1) {
      da = desc_sw128(sa, 4096, 512, 1); db = desc_sw128(sb, 4096, 512, 1);
      idesc = idesc_tf32(128, N, 1, 1);
    } else if (mode == 3) {
      da = desc_sw128(sa, 16, 256, 6); db = desc_sw128(sb, 16, 256, 6);
      idesc = idesc_tf32(128, N, 0, 0);
    } else if (mode == 4) {
      da = desc_sw128(sa, 16, 512, 4); db = desc_sw128(sb, 16, 512, 4);
      idesc = idesc_tf32(128, N, 0, 0);
    } else {
      da = desc_sw128(sa, 16, 1024, 2); db = desc_sw128(sb, 16, 1024, 2);
      idesc = mode == 2 ? idesc_bf16(128, N) : idesc_tf32(128, N, 0, 0);
    }
    t0 = clock64();
    if (mode == 5) {
      // like the 3xTF32 main loop: 4 slots of (A 16 KB | B N*128 B) raw + lo copies, per k-unit 4 (hi,hi) + 4 (hi,lo) + 4 (lo,hi)
      const uint32_t raw = 16384u + (uint32_t)N * 128u;
      const uint32_t idesc5 = idesc_tf32(128, N, 0, 0);
      for (int it = 0; it < niter / 12; ++it) {
        const uint32_t s0 = base + (uint32_t)(it % 2) * 2u * raw;
        const uint64_t a = desc_sw128(s0, 16, 1024, 2), b = desc_sw128(s0 + 16384, 16, 1024, 2);
        const uint64_t lo = raw >> 4;
        for (int pass = 0; pass < 3; ++pass)
          for (int j = 0; j < 4; ++j) {
            const uint64_t aa = a + (pass == 2 ? lo : 0) + 2 * j, bb = b + (pass == 1 ? lo : 0) + 2 * j;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(aa), "l"(bb), "r"(idesc5), "r"(it + pass + j) : "memory");
          }
      }
    } else if (mode == 6 || mode == 7) {
      // mode 6: as mode 5 but A (hi and lo) is read from TMEM columns 384.. (the product kernel's layout)
      // mode 7: mode 6 + a tcgen05.commit per k-unit onto a ring of 2 mbarriers, waiting for the commit of two units ago
      const uint32_t raw = 16384u + (uint32_t)N * 128u;
      const uint32_t idesc5 = idesc_tf32(128, N, 0, 0);
      __shared__ uint64_t ring[2];
      if (mode == 7) { mbar_init(smem_u32(&ring[0]), 1); mbar_init(smem_u32(&ring[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
      for (int it = 0; it < niter / 12; ++it) {
        const uint32_t s0 = base + (uint32_t)(it % 2) * 2u * raw;
        const uint64_t b = desc_sw128(s0 + 16384, 16, 1024, 2);
        const uint64_t lo = raw >> 4;
        const uint32_t a_hi = tmem + 256u + (uint32_t)(it % 2) * 64u, a_lo = a_hi + 32u;
        if (mode == 7 && it >= 2) { while (!mbar_try_wait(smem_u32(&ring[it & 1]), (uint32_t)((it - 2) >> 1) & 1u)) {} }
        for (int pass = 0; pass < 3; ++pass)
          for (int j = 0; j < 4; ++j) {
            const uint32_t aa = (pass == 2 ? a_lo : a_hi) + 8u * j;
            const uint64_t bb = b + (pass == 1 ? lo : 0) + 2 * j;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tmem), "r"(aa), "l"(bb), "r"(idesc5), "r"(it + pass + j) : "memory");
          }
        if (mode == 7)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&ring[it & 1])) : "memory");
      }
    } else
    for (int it = 0; it < niter; ++it) {
      const uint64_t k = (uint64_t)(2 * (it & 3));
      if (mode == 2) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da + k), "l"(db + k), "r"(idesc), "r"(it) : "memory");
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da + (mode == 0 ? k : mode == 4 ? (k & 2) : 0)), "l"(db + (mode == 0 ? k : mode == 4 ? (k & 2) : 0)), "r"(idesc), "r"(it) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    while (!mbar_try_wait(smem_u32(&bar), 0)) {}
    t1 = clock64();
    done = 1;
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else if (hammer == 3 && threadIdx.x >= 32) {
    const uint32_t ta = tmem_base + 448u + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
    uint32_t v = threadIdx.x;
    while (!done) {
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
          ::"r"(ta), "r"(v) : "memory");
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
          ::"r"(ta + 16u), "r"(v) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      ++v;
    }
    if (v == 0x12345u) out[1] = 1;
  } else if (hammer == 4 && threadIdx.x >= 32) {
    // epilogue-like TMEM reads of an accumulator region
    const uint32_t ta = tmem_base + 192u + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
    uint32_t sink = 0;
    while (!done) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(ta) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      sink += r[0] + r[15];
    }
    if (sink == 0x12345u) out[1] = 1;
  } else if (hammer == 1 && threadIdx.x >= 32) {
    float4* reg = reinterpret_cast<float4*>(smem_raw + 1024 + 16384 + 32768);  // 32 KB scratch after A and B
    const int t = threadIdx.x - 32;
    float4 acc = make_float4(0, 0, 0, 0);
    while (!done) {
#pragma unroll 4
      for (int i = t; i < 1024; i += 96) {
        float4 v = reg[i];
        v.x += 1.f; acc.x += v.y;
        reg[i + 1024] = v;
      }
    }
    if (acc.x == 123.f) out[1] = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int niter = 4092 * 16;
  const char* names[8] = {"tf32 K-major SW128", "tf32 MN-major     ", "bf16 K-major SW128", "tf32 K-major SW32 ", "tf32 K-major SW64 ", "tf32 3x pattern   ", "3x, A from TMEM   ", "3x TMEM-A + commit"};
  for (int mode : {7})
    for (int N : {128, 192}) {
      for (int hammer : {0, 1, 3, 4}) {
        const int grid = 148;
        probe<<<grid, 128, 200 * 1024>>>(N, mode, niter, hammer, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long cyc = 0;
        cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        const double per = (double)cyc / niter;
        const int K = mode == 2 ? 16 : 8;
        printf("%s N=%3d hammer=%d: %7.1f cyc/MMA  %6.0f MAC/cyc/SM  (ideal %d cyc)  %s\n", names[mode], N, hammer, per,
               128.0 * N * K / per, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
  return 0;
}
