// tcgen05 / TMA implicit-GEMM convolution for sm_100a: forward, backward-data, backward-filter.
//
// Stride-1 SAME convolutions (every 1x1 / 3x3 / 5x5 Conv2D of the residual cells, the combiners,
// the sampler heads and the pre/post-process towers: SURVEY 8a K1-K3) run as GEMMs whose operands
// are staged by TMA straight from the NHWC fp32 tensors -- there is no im2col buffer and no
// register staging:
//   * the "A" tile of one (tap, 32-channel chunk) is ONE 4-D box {32 ch, tw, th, tn} of the
//     activation tensor at spatial offset (r - pad_t, s - pad_l); TMA zero-fills the halo, which
//     is exactly TF SAME padding.  The box lands in shared memory as 128 rows x 128 B with the
//     128-byte swizzle, i.e. the canonical K-major UMMA operand layout.
//   * forward / dgrad: D[128 pixels, BN] += A[128, 32] * B[BN, 32]^T with kind::tf32, the fp32
//     accumulator lives in TMEM; B is the K-major packed weight copy written by
//     nvae_spectral_norm (forward: [Cout][tap][Cin]; dgrad: the HWIO kernel itself, whose
//     contiguous Cout axis is dgrad's K).
//   * wgrad: the same boxes (loaded with the 32-byte-atom flavour of the 128B swizzle, the only one
//     kind::tf32 takes for MN-major) are consumed as MN-major operands (channels contiguous, pixels = K):
//     D[4 x 32 ci, BN co] += X[pix, ci]^T * dY[pix, co]; each 32-row group of the M tile is an
//     independent (tap, channel-chunk) job so narrow layers (Cin = 32) still fill M = 128.
//     K (= all pixels) is cut over grid.z; partials are reduced in a fixed order (deterministic).
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> bias / residual / accumulate -> 128-bit stores).
// Operands are expected RN-rounded to TF32 by their producers (nvae_bn_act_fwd round_tf32,
// nvae_spectral_norm packs, nvae_round_tf32).
#include <cuda.h>

#include "conv_internal.h"

using namespace nvae;

namespace {

constexpr int kBM = 128;         // UMMA M (TMEM lanes)
constexpr int kChunk = 32;       // fp32 elements per 128-byte swizzle row
constexpr int kMaxStages = 8;
constexpr int kThreads = 192;    // 6 warps
constexpr int kSmemBudget = 220 * 1024;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::tf32, issued by one thread for the CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once all previously issued MMAs have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SWIZZLE_128B)
//   layout 2 = SWIZZLE_128B (16-byte chunks), layout 1 = SWIZZLE_128B_BASE32B (32-byte chunks: the only
//   swizzle kind::tf32 accepts for MN-major operands)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                    uint64_t layout = 2) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (layout << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, M=128, N=n
__host__ __device__ inline uint32_t umma_idesc_tf32(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

struct SmemCtl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full;
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ uint32_t tmem_cols_for(int bn) {
  return bn <= 32 ? 32u : bn <= 64 ? 64u : bn <= 128 ? 128u : 256u;
}

// ------------------------------------------------------------------------------------------------
// forward / backward-data
// ------------------------------------------------------------------------------------------------
struct GemmParams {
  // pixel tiling of the M axis
  int N, H, W;                  // pixel grid (stride 1: identical for input and output)
  int tw, th, tn;               // box extent; rows per tile = tw*th*tn <= 128
  int tiles_h;                  // ceil(H / th) (tn == 1), else 1
  // K loop
  int taps, S;                  // taps = R*S
  int off_h, off_w, dir;        // A box origin shift of tap (r,s): (off_h + dir*r, off_w + dir*s)
  int nchunk1, nchunk2;         // 32-channel chunks of source 1 / source 2
  int k2_base;                  // K index of source 2's first channel inside one tap (= Cin)
  int bk_tap, br_tap;           // B box origin of tap t: (t*bk_tap + k, t*br_tap + n0)
  int BN, stages;
  int passes;                   // 1: TF32;  3: 3xTF32 = (A_hi,B_lo) + (A_lo,B_hi) + (A_hi,B_hi), small terms first
  // epilogue
  int n_valid, n_split;         // columns < n_split -> out1, [n_split, n_valid) -> out2
  float* out1;
  float* out2;
  int ld1, off1, ld2;
  const float* bias;
  const float* res;             // residual laid out like out1
  int accumulate;
};

// index: operand (0 = A source 1, 1 = A source 2, 2 = B) * 2 + plane (0 = value, 1 = low-order TF32 part)
struct TmapSet {
  CUtensorMap m[6];
};

__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ TmapSet maps, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem_raw);
  const uint32_t stage_base = (smem_u32(smem_raw) + (uint32_t)sizeof(SmemCtl) + 1023u) & ~1023u;
  const uint32_t a_bytes = kBM * 128u, b_bytes = (uint32_t)p.BN * 128u, stage_bytes = a_bytes + b_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile -> (n0, h0) ; n tile
  const int mt = blockIdx.x, nt = blockIdx.y;
  int n0, h0;
  if (p.tn > 1) { n0 = mt * p.tn; h0 = 0; }
  else { n0 = mt / p.tiles_h; h0 = (mt - n0 * p.tiles_h) * p.th; }
  const int total = p.passes * p.taps * (p.nchunk1 + p.nchunk2);

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(smem_u32(&ctl->full[i]), 1);
      mbar_init(smem_u32(&ctl->empty[i]), 1);
    }
    mbar_init(smem_u32(&ctl->acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_prefetch_desc(&maps.m[0]);
    tma_prefetch_desc(&maps.m[4]);
  }
  if (warp == 1) tmem_alloc(smem_u32(&ctl->tmem_base), tmem_cols_for(p.BN));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = (uint32_t)(p.tw * p.th * p.tn) * 128u + b_bytes;
      int it = 0;
      for (int pass = 0; pass < p.passes; ++pass) {
        const int a_lo = (p.passes == 3 && pass == 1), b_lo = (p.passes == 3 && pass == 0);
        for (int tap = 0; tap < p.taps; ++tap) {
          const int r = tap / p.S, s = tap - r * p.S;
          const int ah = h0 + p.off_h + p.dir * r, aw = p.off_w + p.dir * s;
          for (int c = 0; c < p.nchunk1 + p.nchunk2; ++c, ++it) {
            const int st = it % p.stages;
            const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
            mbar_wait(smem_u32(&ctl->empty[st]), ph ^ 1u);
            const uint32_t full = smem_u32(&ctl->full[st]);
            const uint32_t sa = stage_base + (uint32_t)st * stage_bytes;
            mbar_expect_tx(full, tx);
            int kk;
            if (c < p.nchunk1) {
              tma_load_4d(sa, &maps.m[a_lo], full, c * kChunk, aw, ah, n0);
              kk = c * kChunk;
            } else {
              tma_load_4d(sa, &maps.m[2 + a_lo], full, (c - p.nchunk1) * kChunk, aw, ah, n0);
              kk = p.k2_base + (c - p.nchunk1) * kChunk;
            }
            tma_load_2d(sa + a_bytes, &maps.m[4 + b_lo], full, tap * p.bk_tap + kk, tap * p.br_tap + nt * p.BN);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(p.BN, 0, 0);
      for (int it = 0; it < total; ++it) {
        const int st = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(smem_u32(&ctl->full[st]), ph);
        tc_fence_after();
        const uint32_t sa = stage_base + (uint32_t)st * stage_bytes;
        const uint64_t da = umma_desc_sw128(sa, 16, 1024), db = umma_desc_sw128(sa + a_bytes, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 4 x (K = 8 tf32 = 32 bytes) inside the 128-byte swizzle row
          umma_tf32(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it | k) != 0);
        umma_commit(smem_u32(&ctl->empty[st]));
      }
      umma_commit(smem_u32(&ctl->acc_full));
    }
  } else {
    // epilogue: warp w owns TMEM lanes [32*(w%4), +32) = tile rows
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    const int per_img = p.tw * p.th;
    const int in = row / per_img, rem = row - in * per_img;
    const int ih = rem / p.tw, iw = rem - ih * p.tw;
    const bool row_ok = row < per_img * p.tn && (n0 + in) < p.N && (h0 + ih) < p.H;
    const int64_t pix = ((int64_t)(n0 + in) * p.H + (h0 + ih)) * p.W + iw;
    mbar_wait(smem_u32(&ctl->acc_full), 0);
    tc_fence_after();
    const int ncol0 = nt * p.BN;
    for (int j = 0; j < p.BN; j += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)j, v);
      if (!row_ok) continue;
#pragma unroll
      for (int q = 0; q < 32; q += 4) {
        const int n = ncol0 + j + q;
        if (j + q >= p.BN || n >= p.n_valid) break;
        float4 o = make_float4(__uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]),
                               __uint_as_float(v[q + 3]));
        float* dst;
        if (n < p.n_split) {
          dst = p.out1 + pix * p.ld1 + p.off1 + n;
          if (p.bias) {
            const float4 b = ldg4(p.bias + n);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          if (p.res) {
            const float4 b = ldg4(p.res + pix * p.ld1 + p.off1 + n);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
        } else {
          if (p.out2 == nullptr) continue;
          dst = p.out2 + pix * p.ld2 + (n - p.n_split);
        }
        if (p.accumulate) {
          const float4 b = *reinterpret_cast<const float4*>(dst);
          o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        }
        stg4(dst, o);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, tmem_cols_for(p.BN));
  }
}

// ------------------------------------------------------------------------------------------------
// backward-filter
// ------------------------------------------------------------------------------------------------
struct WgradParams {
  int N, H, W;
  int tw, th, tn, KP;           // pixel box; KP = tw*th*tn (multiple of 8) = K extent of one stage
  int tiles_h, n_ptiles;        // pixel tiles per image column / total
  int ptiles_per_split;
  int S, pad_t, pad_l;
  int nchunk1, nchunk2, njobs;  // job = (tap, 32-channel chunk); 4 jobs per M tile
  int Cin, Cin2, Ct, Cout;
  int BN, stages;
  int passes;                   // 1 or 3 (x_hi*dy_lo, x_lo*dy_hi, x_hi*dy_hi)
  float* out;                   // dw (HWIO) or the [splits][taps*Ct][Cout] partial buffer
  int64_t split_stride;         // elements between partials (0 when not split)
};

// maps: x source 1 {value, lo}, x source 2 {value, lo}, dy {value, lo}
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_tc_kernel(const __grid_constant__ TmapSet maps, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem_raw);
  const uint32_t stage_base = (smem_u32(smem_raw) + (uint32_t)sizeof(SmemCtl) + 1023u) & ~1023u;
  const uint32_t box_bytes = (uint32_t)p.KP * 128u;
  const int nb = p.BN / kChunk;
  const uint32_t a_bytes = 4u * box_bytes, stage_bytes = a_bytes + (uint32_t)nb * box_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, nt = blockIdx.y, split = blockIdx.z;
  const int nch = p.nchunk1 + p.nchunk2;
  const int pt0 = split * p.ptiles_per_split;
  const int pt1 = min(pt0 + p.ptiles_per_split, p.n_ptiles);
  const int npt = pt1 - pt0;
  const int total = npt * p.passes;
  int njob = p.njobs - mt * 4;
  njob = njob > 4 ? 4 : njob;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(smem_u32(&ctl->full[i]), 1);
      mbar_init(smem_u32(&ctl->empty[i]), 1);
    }
    mbar_init(smem_u32(&ctl->acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_prefetch_desc(&maps.m[0]);
    tma_prefetch_desc(&maps.m[4]);
  }
  if (warp == 1) tmem_alloc(smem_u32(&ctl->tmem_base), tmem_cols_for(p.BN));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      // the (tap, chunk) of each of this tile's jobs
      int jc[4], jh[4], jw[4], js[4];
      for (int j = 0; j < 4; ++j) {
        const int job = mt * 4 + j;
        const int tap = job / nch, c = job - tap * nch;
        const int r = tap / p.S, s = tap - r * p.S;
        jh[j] = r - p.pad_t; jw[j] = s - p.pad_l;
        js[j] = c >= p.nchunk1;
        jc[j] = (js[j] ? c - p.nchunk1 : c) * kChunk;
      }
      const uint32_t tx = (uint32_t)(njob + nb) * box_bytes;
      for (int it = 0; it < total; ++it) {
        const int pass = it / npt;
        const int x_lo = (p.passes == 3 && pass == 1), dy_lo = (p.passes == 3 && pass == 0);
        const int pt = pt0 + (it - pass * npt);
        int n0, h0;
        if (p.tn > 1) { n0 = pt * p.tn; h0 = 0; }
        else { n0 = pt / p.tiles_h; h0 = (pt - n0 * p.tiles_h) * p.th; }
        const int st = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(smem_u32(&ctl->empty[st]), ph ^ 1u);
        const uint32_t full = smem_u32(&ctl->full[st]);
        const uint32_t sa = stage_base + (uint32_t)st * stage_bytes;
        mbar_expect_tx(full, tx);
        for (int j = 0; j < njob; ++j)
          tma_load_4d(sa + (uint32_t)j * box_bytes, &maps.m[2 * js[j] + x_lo], full, jc[j], jw[j], h0 + jh[j], n0);
        for (int b = 0; b < nb; ++b)
          tma_load_4d(sa + a_bytes + (uint32_t)b * box_bytes, &maps.m[4 + dy_lo], full, nt * p.BN + b * kChunk, 0, h0,
                      n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(p.BN, 1, 1);
      const int ksteps = p.KP / 8;
      for (int it = 0; it < total; ++it) {
        const int st = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(smem_u32(&ctl->full[st]), ph);
        tc_fence_after();
        const uint32_t sa = stage_base + (uint32_t)st * stage_bytes;
        // MN-major, 128B swizzle with 32B atoms: 32 channels x 4 pixels per 512-byte atom;
        // LBO = next 32-channel box, SBO = next 4 pixels; one K=8 MMA spans two atoms
        const uint64_t da = umma_desc_sw128(sa, box_bytes, 512, 1), db = umma_desc_sw128(sa + a_bytes, box_bytes, 512, 1);
        for (int k = 0; k < ksteps; ++k)
          umma_tf32(tmem, da + (uint64_t)(64 * k), db + (uint64_t)(64 * k), idesc, (it | k) != 0);
        umma_commit(smem_u32(&ctl->empty[st]));
      }
      umma_commit(smem_u32(&ctl->acc_full));
    }
  } else {
    const int lg = warp & 3;  // == job index inside the tile: TMEM lanes [32*lg, +32) are its 32 input channels
    const int job = mt * 4 + lg;
    const int tap = job / nch, c = job - tap * nch;
    const bool s2 = c >= p.nchunk1;
    const int ch = (s2 ? c - p.nchunk1 : c) * kChunk + lane;
    const bool row_ok = total > 0 && job < p.njobs && ch < (s2 ? p.Cin2 : p.Cin);
    const int64_t wrow = (int64_t)tap * p.Ct + (s2 ? p.Cin : 0) + ch;
    float* dst_row = p.out + (int64_t)split * p.split_stride + wrow * p.Cout;
    if (total > 0) {
      mbar_wait(smem_u32(&ctl->acc_full), 0);
      tc_fence_after();
    }
    const int ncol0 = nt * p.BN;
    for (int j = 0; j < p.BN; j += 32) {
      uint32_t v[32];
      if (total > 0) tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)j, v);
      if (!(job < p.njobs && ch < (s2 ? p.Cin2 : p.Cin))) continue;
#pragma unroll
      for (int q = 0; q < 32; q += 4) {
        const int n = ncol0 + j + q;
        if (n >= p.Cout) break;
        float4 o = row_ok ? make_float4(__uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]),
                                        __uint_as_float(v[q + 3]))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
        stg4(dst_row + n, o);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, tmem_cols_for(p.BN));
  }
}

__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ part, int64_t n4, int splits, float* __restrict__ dw) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = ldg4(part + 4 * i);
    for (int k = 1; k < splits; ++k) {
      const float4 t = ldg4(part + (int64_t)k * n4 * 4 + 4 * i);
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    stg4(dw + 4 * i, s);
  }
}

__global__ void round_tf32_kernel(float* __restrict__ p, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = *reinterpret_cast<float4*>(p + 4 * i);
    v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
    stg4(p + 4 * i, v);
  }
}

// lo = tf32_rn(v - trunc_tf32(v)): kind::tf32 reads only the top 19 bits of v (truncation), so v ~= trunc(v) + lo
// to ~2^-21 relative and  a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi  (3xTF32, fp32-level products).
// v is a [rows][C] view with leading dimension ld; lo is compact [rows][C].
__global__ void tf32_split_kernel(const float* __restrict__ v, int64_t n4, int C4, int ld, float* __restrict__ lo) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C4;
    const int c4 = (int)(i - row * C4);
    const float4 a = ldg4(v + row * ld + 4 * c4);
    float4 o;
    o.x = round_tf32(a.x - __uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u));
    o.y = round_tf32(a.y - __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u));
    o.z = round_tf32(a.z - __uint_as_float(__float_as_uint(a.z) & 0xFFFFE000u));
    o.w = round_tf32(a.w - __uint_as_float(__float_as_uint(a.w) & 0xFFFFE000u));
    stg4(lo + 4 * i, o);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// 4-D map over an NHWC fp32 tensor view: channels [c_off, c_off + C) of rows with leading dimension ld
int make_map_nhwc(CUtensorMap* m, const float* base, int N, int H, int W, int C, int ld, int tw, int th, int tn,
                  CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return NVAE_E_DRIVER;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)W * ld * 4, (cuuint64_t)H * W * ld * 4};
  cuuint32_t box[4] = {(cuuint32_t)kChunk, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NVAE_OK : NVAE_E_DRIVER;
}
// 2-D map over a row-major [rows][cols] fp32 matrix, box {32 cols, box_rows}
int make_map_2d(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return NVAE_E_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {(cuuint32_t)kChunk, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NVAE_OK : NVAE_E_DRIVER;
}

struct PixTile { int tw, th, tn, tiles_h, n_tiles; };

// rows per tile <= max_rows; rows % row_mult == 0 (rows beyond H are TMA zero fill)
bool pick_pix_tile(int N, int H, int W, int max_rows, int row_mult, PixTile* t) {
  if (W > max_rows || W > 256) return false;
  int th = max_rows / W;
  if (th > H) th = H;
  int tiles_h = (H + th - 1) / th;
  th = (H + tiles_h - 1) / tiles_h;  // even out
  while ((W * th) % row_mult != 0) {  // pad with rows that fall outside the image
    ++th;
    if (W * th > max_rows || th > 256) return false;
  }
  tiles_h = (H + th - 1) / th;
  int tn = 1;
  if (tiles_h == 1) {
    tn = max_rows / (W * th);
    if (tn > N) tn = N;
    if (tn < 1) tn = 1;
    if (tn > 256) tn = 256;
  }
  t->tw = W; t->th = th; t->tn = tn; t->tiles_h = tiles_h;
  t->n_tiles = tn > 1 ? (N + tn - 1) / tn : N * tiles_h;
  return true;
}

int pick_bn(int n_total, int m_tiles, int mult) {
  int nt = (n_total + 255) / 256;
  int bn = (int)round_up(ceil_div(n_total, nt), mult);
  while ((int64_t)m_tiles * nt < kNumSMs && bn > 64) {
    nt *= 2;
    bn = (int)round_up(ceil_div(n_total, nt), mult);
  }
  return bn;
}

int pick_stages(size_t stage_bytes) {
  int s = (int)((kSmemBudget - 2048) / stage_bytes);
  return s > kMaxStages ? kMaxStages : s;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

bool common_ok(const NvaeConvDesc* d) {
  if (d->stride != 1 || d->Ho != d->H || d->Wo != d->W) return false;
  if (d->W > 128) return false;
  if ((d->Cin & 3) || (d->Cin2 & 3) || (d->Cout & 3) || d->Cout < 8) return false;
  if (d->Cin2 > 0 && (d->Cin % kChunk) != 0) return false;
  if ((d->y_ld & 3) || (d->y_off & 3)) return false;
  if (!(d->pre_scale == 0.f && d->pre_shift == 0.f) && !(d->pre_scale == 1.f && d->pre_shift == 0.f)) return false;
  return true;
}

struct WgradPlan { PixTile t; int BN, stages, splits, ptiles_per_split, njobs, m_tiles, n_tiles; };

bool plan_wgrad(const NvaeConvDesc* d, WgradPlan* w) {
  if (!pick_pix_tile(d->N, d->H, d->W, 32, 8, &w->t) && !pick_pix_tile(d->N, d->H, d->W, 64, 8, &w->t) &&
      !pick_pix_tile(d->N, d->H, d->W, 128, 8, &w->t))
    return false;
  const int KP = w->t.tw * w->t.th * w->t.tn;
  const int nch = (d->Cin + kChunk - 1) / kChunk + (d->Cin2 + kChunk - 1) / kChunk;
  w->njobs = d->R * d->S * nch;
  w->m_tiles = (w->njobs + 3) / 4;
  w->BN = pick_bn(d->Cout, w->m_tiles, 32);
  // keep at least 3 stages in shared memory
  while ((size_t)(4 + w->BN / kChunk) * KP * 128 * 3 > (size_t)kSmemBudget - 2048 && w->BN > 32) w->BN -= 32;
  if ((size_t)(4 + w->BN / kChunk) * KP * 128 * 2 > (size_t)kSmemBudget - 2048) return false;
  w->n_tiles = (d->Cout + w->BN - 1) / w->BN;
  w->stages = pick_stages((size_t)(4 + w->BN / kChunk) * KP * 128);
  const int64_t tiles = (int64_t)w->m_tiles * w->n_tiles;
  int64_t splits = ceil_div(2 * kNumSMs, tiles);
  const int64_t max_splits = ceil_div(w->t.n_tiles, 8);  // at least 8 pixel tiles per CTA
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  w->ptiles_per_split = (int)ceil_div(w->t.n_tiles, splits);
  w->splits = (int)ceil_div(w->t.n_tiles, w->ptiles_per_split);
  return true;
}

template <class K>
int set_smem_attr(K kernel) {
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return (int)e;
    done = true;
  }
  return NVAE_OK;
}

size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

// Workspace layout (bytes): [wgrad split-K partials][x lo][x2 lo][dy lo]; the lo planes exist in 3xTF32 mode only.
struct WsPlan { size_t part, x_lo, x2_lo, dy_lo, total; };

WsPlan plan_ws(const NvaeConvDesc* d, int which, int splits) {
  WsPlan w{};
  const bool x3 = d->precision == NVAE_PREC_TF32X3;
  const size_t pix = (size_t)d->N * d->H * d->W;
  size_t off = 0;
  w.part = off;
  if (which == 2 && splits > 1) off += al256((size_t)splits * d->R * d->S * (d->Cin + d->Cin2) * d->Cout * sizeof(float));
  w.x_lo = off;
  if (x3 && which != 1) off += al256(pix * d->Cin * sizeof(float));
  w.x2_lo = off;
  if (x3 && which != 1) off += al256(pix * d->Cin2 * sizeof(float));
  w.dy_lo = off;
  if (x3 && which != 0) off += al256(pix * d->Cout * sizeof(float));
  w.total = off;
  return w;
}

int split_lo(const float* v, int64_t rows, int C, int ld, float* lo, cudaStream_t stream) {
  const int64_t n4 = rows * (C / 4);
  if (n4 <= 0) return NVAE_OK;
  int64_t g = ceil_div(n4, 256);
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  tf32_split_kernel<<<(int)g, 256, 0, stream>>>(v, n4, C / 4, ld, lo);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

// plane 1 (low-order parts) of a packed weight copy follows plane 0 at this element offset
int64_t weight_plane(const NvaeConvDesc* d) {
  return round_up((int64_t)d->R * d->S * (d->Cin + d->Cin2) * d->Cout, 4);
}

}  // namespace

bool nvae_conv_tc_supported(const NvaeConvDesc* d, int which) {
  if (!common_ok(d)) return false;
  PixTile t;
  if (which == 0 || which == 1) return pick_pix_tile(d->N, d->H, d->W, kBM, 1, &t);
  WgradPlan w;
  return plan_wgrad(d, &w);
}

size_t nvae_conv_tc_ws_bytes(const NvaeConvDesc* d, int which) {
  if (!common_ok(d)) return 0;
  int splits = 1;
  if (which == 2) {
    WgradPlan w;
    if (!plan_wgrad(d, &w)) return 0;
    splits = w.splits;
  }
  return plan_ws(d, which, splits).total;
}

int nvae_conv2d_fwd_tc(const NvaeConvDesc* d, const float* x, const float* x2, const float* w_tr, const float* bias,
                       const float* residual, float* y, void* ws, size_t ws_bytes, cudaStream_t stream) {
  PixTile t;
  if (!common_ok(d) || !pick_pix_tile(d->N, d->H, d->W, kBM, 1, &t)) return NVAE_E_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(x2) || !aligned16(w_tr) || !aligned16(bias) || !aligned16(residual) || !aligned16(y) ||
      !aligned16(ws))
    return NVAE_E_UNSUPPORTED;
  const bool x3 = d->precision == NVAE_PREC_TF32X3;
  const int Ct = d->Cin + d->Cin2, taps = d->R * d->S;
  const int64_t pix = (int64_t)d->N * d->H * d->W;
  GemmParams p{};
  p.N = d->N; p.H = d->H; p.W = d->W;
  p.tw = t.tw; p.th = t.th; p.tn = t.tn; p.tiles_h = t.tiles_h;
  p.taps = taps; p.S = d->S;
  p.off_h = -d->pad_t; p.off_w = -d->pad_l; p.dir = 1;
  p.nchunk1 = (d->Cin + kChunk - 1) / kChunk;
  p.nchunk2 = (d->Cin2 + kChunk - 1) / kChunk;
  p.k2_base = d->Cin;
  p.bk_tap = Ct; p.br_tap = 0;
  p.BN = pick_bn(d->Cout, t.n_tiles, 16);
  p.stages = pick_stages((size_t)kBM * 128 + (size_t)p.BN * 128);
  p.passes = x3 ? 3 : 1;
  p.n_valid = d->Cout; p.n_split = d->Cout;
  p.out1 = y; p.out2 = nullptr;
  p.ld1 = d->y_ld > 0 ? d->y_ld : d->Cout; p.off1 = d->y_off; p.ld2 = 0;
  p.bias = bias; p.res = residual; p.accumulate = 0;
  const float *x_lo = x, *x2_lo = x2, *w_lo = w_tr;
  if (x3) {
    const WsPlan wp = plan_ws(d, 0, 1);
    if (ws == nullptr || ws_bytes < wp.total) return NVAE_E_WORKSPACE;
    float* l1 = reinterpret_cast<float*>((char*)ws + wp.x_lo);
    float* l2 = reinterpret_cast<float*>((char*)ws + wp.x2_lo);
    int rc = split_lo(x, pix, d->Cin, d->Cin, l1, stream);
    if (rc) return rc;
    if (d->Cin2 > 0 && (rc = split_lo(x2, pix, d->Cin2, d->Cin2, l2, stream))) return rc;
    x_lo = l1; x2_lo = l2;
    w_lo = w_tr + weight_plane(d);
  }
  TmapSet maps;
  int rc = make_map_nhwc(&maps.m[0], x, d->N, d->H, d->W, d->Cin, d->Cin, t.tw, t.th, t.tn);
  if (rc) return rc;
  rc = make_map_nhwc(&maps.m[1], x_lo, d->N, d->H, d->W, d->Cin, d->Cin, t.tw, t.th, t.tn);
  if (rc) return rc;
  if (d->Cin2 > 0) {
    rc = make_map_nhwc(&maps.m[2], x2, d->N, d->H, d->W, d->Cin2, d->Cin2, t.tw, t.th, t.tn);
    if (rc) return rc;
    rc = make_map_nhwc(&maps.m[3], x2_lo, d->N, d->H, d->W, d->Cin2, d->Cin2, t.tw, t.th, t.tn);
    if (rc) return rc;
  } else {
    maps.m[2] = maps.m[0]; maps.m[3] = maps.m[1];
  }
  rc = make_map_2d(&maps.m[4], w_tr, d->Cout, (int64_t)taps * Ct, p.BN);
  if (rc) return rc;
  rc = make_map_2d(&maps.m[5], w_lo, d->Cout, (int64_t)taps * Ct, p.BN);
  if (rc) return rc;
  rc = set_smem_attr(conv_gemm_tc_kernel);
  if (rc) return rc;
  const size_t smem = sizeof(SmemCtl) + 1024 + (size_t)p.stages * ((size_t)kBM * 128 + (size_t)p.BN * 128);
  dim3 grid((unsigned)t.n_tiles, (unsigned)ceil_div(d->Cout, p.BN), 1);
  conv_gemm_tc_kernel<<<grid, kThreads, smem, stream>>>(maps, p);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

int nvae_conv2d_dgrad_tc(const NvaeConvDesc* d, const float* dy, const float* w_rnd, float* dx, float* dx2,
                         int accumulate, void* ws, size_t ws_bytes, cudaStream_t stream) {
  PixTile t;
  if (!common_ok(d) || !pick_pix_tile(d->N, d->H, d->W, kBM, 1, &t)) return NVAE_E_UNSUPPORTED;
  if (!aligned16(dy) || !aligned16(w_rnd) || !aligned16(dx) || !aligned16(dx2) || !aligned16(ws))
    return NVAE_E_UNSUPPORTED;
  if (dx == nullptr) return NVAE_E_UNSUPPORTED;
  const bool x3 = d->precision == NVAE_PREC_TF32X3;
  const int Ct = d->Cin + d->Cin2, taps = d->R * d->S;
  const int ld = d->y_ld > 0 ? d->y_ld : d->Cout;
  const int64_t pix = (int64_t)d->N * d->H * d->W;
  GemmParams p{};
  p.N = d->N; p.H = d->H; p.W = d->W;
  p.tw = t.tw; p.th = t.th; p.tn = t.tn; p.tiles_h = t.tiles_h;
  p.taps = taps; p.S = d->S;
  p.off_h = d->pad_t; p.off_w = d->pad_l; p.dir = -1;  // dx[h,w] = sum dy[h + pad_t - r, w + pad_l - s] * w[r,s]
  p.nchunk1 = (d->Cout + kChunk - 1) / kChunk;
  p.nchunk2 = 0;
  p.k2_base = 0;
  p.bk_tap = 0; p.br_tap = Ct;
  p.BN = pick_bn(Ct, t.n_tiles, 16);
  p.stages = pick_stages((size_t)kBM * 128 + (size_t)p.BN * 128);
  p.passes = x3 ? 3 : 1;
  p.n_valid = Ct; p.n_split = d->Cin;
  p.out1 = dx; p.out2 = dx2;
  p.ld1 = d->Cin; p.off1 = 0; p.ld2 = d->Cin2;
  p.bias = nullptr; p.res = nullptr; p.accumulate = accumulate;
  TmapSet maps;
  int rc = make_map_nhwc(&maps.m[0], dy + d->y_off, d->N, d->H, d->W, d->Cout, ld, t.tw, t.th, t.tn);
  if (rc) return rc;
  const float* w_lo = w_rnd;
  if (x3) {
    const WsPlan wp = plan_ws(d, 1, 1);
    if (ws == nullptr || ws_bytes < wp.total) return NVAE_E_WORKSPACE;
    float* l = reinterpret_cast<float*>((char*)ws + wp.dy_lo);
    rc = split_lo(dy + d->y_off, pix, d->Cout, ld, l, stream);
    if (rc) return rc;
    rc = make_map_nhwc(&maps.m[1], l, d->N, d->H, d->W, d->Cout, d->Cout, t.tw, t.th, t.tn);
    if (rc) return rc;
    w_lo = w_rnd + weight_plane(d);
  } else {
    maps.m[1] = maps.m[0];
  }
  maps.m[2] = maps.m[0]; maps.m[3] = maps.m[1];
  rc = make_map_2d(&maps.m[4], w_rnd, (int64_t)taps * Ct, d->Cout, p.BN);
  if (rc) return rc;
  rc = make_map_2d(&maps.m[5], w_lo, (int64_t)taps * Ct, d->Cout, p.BN);
  if (rc) return rc;
  rc = set_smem_attr(conv_gemm_tc_kernel);
  if (rc) return rc;
  const size_t smem = sizeof(SmemCtl) + 1024 + (size_t)p.stages * ((size_t)kBM * 128 + (size_t)p.BN * 128);
  dim3 grid((unsigned)t.n_tiles, (unsigned)ceil_div(Ct, p.BN), 1);
  conv_gemm_tc_kernel<<<grid, kThreads, smem, stream>>>(maps, p);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}

int nvae_conv2d_wgrad_tc(const NvaeConvDesc* d, const float* x, const float* x2, const float* dy, float* dw, void* ws,
                         size_t ws_bytes, cudaStream_t stream) {
  WgradPlan w;
  if (!common_ok(d) || !plan_wgrad(d, &w)) return NVAE_E_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(x2) || !aligned16(dy) || !aligned16(dw) || !aligned16(ws)) return NVAE_E_UNSUPPORTED;
  const bool x3 = d->precision == NVAE_PREC_TF32X3;
  const int Ct = d->Cin + d->Cin2, taps = d->R * d->S;
  const int ld = d->y_ld > 0 ? d->y_ld : d->Cout;
  const int64_t wsize = (int64_t)taps * Ct * d->Cout;
  const int64_t pix = (int64_t)d->N * d->H * d->W;
  const WsPlan wp = plan_ws(d, 2, w.splits);
  if (wp.total > 0 && (ws == nullptr || ws_bytes < wp.total)) return NVAE_E_WORKSPACE;
  WgradParams p{};
  p.N = d->N; p.H = d->H; p.W = d->W;
  p.tw = w.t.tw; p.th = w.t.th; p.tn = w.t.tn; p.KP = w.t.tw * w.t.th * w.t.tn;
  p.tiles_h = w.t.tiles_h; p.n_ptiles = w.t.n_tiles; p.ptiles_per_split = w.ptiles_per_split;
  p.S = d->S; p.pad_t = d->pad_t; p.pad_l = d->pad_l;
  p.nchunk1 = (d->Cin + kChunk - 1) / kChunk;
  p.nchunk2 = (d->Cin2 + kChunk - 1) / kChunk;
  p.njobs = w.njobs;
  p.Cin = d->Cin; p.Cin2 = d->Cin2; p.Ct = Ct; p.Cout = d->Cout;
  p.BN = w.BN; p.stages = w.stages;
  p.passes = x3 ? 3 : 1;
  if (w.splits > 1) {
    p.out = reinterpret_cast<float*>((char*)ws + wp.part);
    p.split_stride = wsize;
  } else {
    p.out = dw;
    p.split_stride = 0;
  }
  const float *x_lo = x, *x2_lo = x2, *dy_hi = dy + d->y_off, *dy_lo = dy_hi;
  int dy_lo_ld = ld;
  int rc;
  if (x3) {
    float* l1 = reinterpret_cast<float*>((char*)ws + wp.x_lo);
    float* l2 = reinterpret_cast<float*>((char*)ws + wp.x2_lo);
    float* l3 = reinterpret_cast<float*>((char*)ws + wp.dy_lo);
    if ((rc = split_lo(x, pix, d->Cin, d->Cin, l1, stream))) return rc;
    if (d->Cin2 > 0 && (rc = split_lo(x2, pix, d->Cin2, d->Cin2, l2, stream))) return rc;
    if ((rc = split_lo(dy_hi, pix, d->Cout, ld, l3, stream))) return rc;
    x_lo = l1; x2_lo = l2; dy_lo = l3; dy_lo_ld = d->Cout;
  }
  TmapSet maps;
  const CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  rc = make_map_nhwc(&maps.m[0], x, d->N, d->H, d->W, d->Cin, d->Cin, p.tw, p.th, p.tn, swz);
  if (rc) return rc;
  rc = make_map_nhwc(&maps.m[1], x_lo, d->N, d->H, d->W, d->Cin, d->Cin, p.tw, p.th, p.tn, swz);
  if (rc) return rc;
  if (d->Cin2 > 0) {
    rc = make_map_nhwc(&maps.m[2], x2, d->N, d->H, d->W, d->Cin2, d->Cin2, p.tw, p.th, p.tn, swz);
    if (rc) return rc;
    rc = make_map_nhwc(&maps.m[3], x2_lo, d->N, d->H, d->W, d->Cin2, d->Cin2, p.tw, p.th, p.tn, swz);
    if (rc) return rc;
  } else {
    maps.m[2] = maps.m[0]; maps.m[3] = maps.m[1];
  }
  rc = make_map_nhwc(&maps.m[4], dy_hi, d->N, d->H, d->W, d->Cout, ld, p.tw, p.th, p.tn, swz);
  if (rc) return rc;
  rc = make_map_nhwc(&maps.m[5], dy_lo, d->N, d->H, d->W, d->Cout, dy_lo_ld, p.tw, p.th, p.tn, swz);
  if (rc) return rc;
  rc = set_smem_attr(conv_wgrad_tc_kernel);
  if (rc) return rc;
  const size_t smem = sizeof(SmemCtl) + 1024 + (size_t)p.stages * (size_t)(4 + p.BN / kChunk) * p.KP * 128;
  dim3 grid((unsigned)w.m_tiles, (unsigned)w.n_tiles, (unsigned)w.splits);
  conv_wgrad_tc_kernel<<<grid, kThreads, smem, stream>>>(maps, p);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  if (w.splits > 1) {
    int64_t g = ceil_div(wsize / 4, 256);
    if (g > kNumSMs * 8) g = kNumSMs * 8;
    wgrad_tc_reduce_kernel<<<(int)g, 256, 0, stream>>>(p.out, wsize / 4, w.splits, dw);
    NVAE_RETURN_IF_LAUNCH_FAILED();
  }
  return NVAE_OK;
}

int nvae_round_tf32_inplace(float* p, int64_t n, cudaStream_t stream) {
  if (n <= 0) return NVAE_OK;
  if ((n & 3) || !aligned16(p)) return NVAE_E_BADSHAPE;
  int64_t g = ceil_div(n / 4, 256);
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  round_tf32_kernel<<<(int)g, 256, 0, stream>>>(p, n / 4);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}
