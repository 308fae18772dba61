// tcgen05 / TMA implicit-GEMM convolution for sm_100a: forward, backward-data, backward-filter.
//
// Stride-1 SAME convolutions (every 1x1 / 3x3 / 5x5 Conv2D of the residual cells, the combiners,
// the sampler heads and the pre/post-process towers: SURVEY 8a K1-K3) run as GEMMs whose operands
// are staged by TMA straight from the NHWC fp32 tensors -- there is no im2col buffer:
//   * the "A" tile of one (tap, 32-channel chunk) is ONE 4-D box {32 ch, tw, th, tn} of the
//     activation tensor at spatial offset (r - pad_t, s - pad_l); TMA zero-fills the halo, which
//     is exactly TF SAME padding.  The box lands in shared memory as 128 rows x 128 B with the
//     128-byte swizzle, i.e. the canonical K-major UMMA operand layout.
//   * forward / dgrad: D[128 pixels, BN] += A[128, 32] * B[BN, 32]^T with kind::tf32, the fp32
//     accumulator lives in TMEM; B is the K-major weight copy written by nvae_spectral_norm
//     (forward: [Cout][tap][Cin]; dgrad: the HWIO layout, whose contiguous Cout axis is dgrad's K).
//   * wgrad: the same boxes (loaded with the 32-byte-atom flavour of the 128B swizzle, the only one
//     kind::tf32 takes for MN-major) are consumed as MN-major operands (channels contiguous, pixels = K):
//     D[4 x 32 ci, BN co] += X[pix, ci]^T * dY[pix, co]; each 32-row group of the M tile is an
//     independent (tap, channel-chunk) job so narrow layers (Cin = 32) still fill M = 128.
// Arithmetic: NVAE_PREC_TF32 issues one MMA per operand pair.  NVAE_PREC_TF32X3 (fp32-level accuracy, the
// default) keeps the raw fp32 tiles as the high parts (kind::tf32 reads the top 19 bits, i.e. truncates),
// has eight converter warps write the TF32-rounded remainders  lo = rn(v - trunc(v))  of every staged tile
// to a second shared-memory ring, and issues three MMAs per K step: A_hi*B_hi as soon as the tile lands (the
// converters run meanwhile), then A_hi*B_lo + A_lo*B_hi.  Nothing low-order ever touches HBM or L2.
// The large GEMMs of that mode (>= 20 GFLOP per launch: the six 5x5 convolutions of the postprocess tower, 87 % of the
// step's FLOPs) run the same three-term product as 3xFP16 on kind::f16 instead (TcParams::f16): operands scaled by a power
// of two from their absmax, A split by the converters into packed fp16 pairs in TMEM, B (weights / dY) split ahead of
// the launch into rows the SAME tensor maps stage, two accumulators per CTA sharing the staged tiles (nsub / dual) --
// twice the MMA rate, half the shared-memory bytes per stage.  DESIGN.md 3.1b.
// Scheduling: stream-K.  The (tile, k-unit) space is cut into gridDim.x equal contiguous ranges (one CTA per
// SM, at most 148), so every SM gets the same number of pipeline stages whatever the tile count; a CTA
// that covers only part of a tile's K range writes its raw accumulator to a partial buffer and a fix-up
// kernel sums the partials of each split tile in CTA order (deterministic) and applies the epilogue.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer (warp 14: a second issuer that takes
// the second accumulator of the two-accumulator 3xFP16 tiles -- the issuing thread, not the tensor pipe, paces those kernels),
// warps 2..9 = lo-part converters, warps 10..13 = epilogue (tcgen05.ld -> bias / residual / accumulate ->
// 128-bit stores) on a double-buffered TMEM accumulator, so the next tile's MMAs overlap the drain.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "conv_internal.h"

using namespace nvae;

namespace {

constexpr int kBM = 128;         // UMMA M (TMEM lanes)
constexpr int kChunk = 32;       // fp32 elements per 128-byte swizzle row
constexpr int kMaxStages = 8;
constexpr int kConvThreads = 256;                   // warps 2..9: lo-part converters (3xTF32)
constexpr int kEpiThreads = 128;                    // warps 10..13: epilogue, one per TMEM lane group
constexpr int kIssue2Warp = (64 + kConvThreads + kEpiThreads) / 32;  // warp 14: second MMA issuer (two-accumulator 3xFP16 tiles)
constexpr int kThreads = 64 + kConvThreads + kEpiThreads + 32;
constexpr int kSmemBudget = 208 * 1024;  // operand rings; the epilogue staging (17 KB) and barriers come on top

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
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
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
// same with the A operand read from TMEM (lanes = rows, one 32-bit column per K element)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 (fp16 operands, fp32 accumulate), A from TMEM: lanes = rows, each 32-bit column holds two consecutive K
// elements (even k in the low half); one instruction covers K = 16, i.e. 8 columns
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// cta_group::2 forms: ONE instruction, issued by the leader CTA of a 2-CTA cluster, drives both SMs' tensor cores --
// D[256 x N]: each CTA's TMEM holds its 128 rows of A and of D, each CTA's shared memory holds N/2 rows of B.
__device__ __forceinline__ void umma_tf32_ts_2cta(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once all previously issued MMAs completed) on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// arrive on the barrier at this shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank) : "memory");
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
__host__ __device__ inline uint32_t umma_idesc_tf32(int n, int a_mn_major, int b_mn_major, int m = kBM) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D=F32, A=B=F16 (format 0), K-major A (TMEM), M=128, N=n
__host__ __device__ inline uint32_t umma_idesc_f16(int n, int b_mn_major, int m = kBM) {
  return (1u << 4) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

#ifdef NVAE_TC_TIMING
// development build only: per-CTA timestamps (ns, %globaltimer) [start, setup done, first tile landed, last MMA
// issued, accumulator complete, epilogue done, exit]
__device__ unsigned long long g_tc_timing[148 * 16];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP(i) g_tc_timing[blockIdx.x * 16 + (i)] = gtime()
__device__ long long g_tc_cyc[8 * 16];
#define TC_CYC(it, j) do { if (blockIdx.x == 0 && (it) >= 8 && (it) < 16) g_tc_cyc[((it) - 8) * 16 + (j)] = clock64(); } while (0)
// whole-launch totals for CTA 0, issuer 0: cycles waiting on full[] (TMA), on conv[] (converters), total loop cycles, units
__device__ long long g_tc_wait[8];
#define TC_WAIT_BEGIN() long long tcw_t0 = clock64(), tcw_a = 0, tcw_full = 0, tcw_conv = 0, tcw_n = 0
#define TC_WAIT_A() tcw_a = clock64()
#define TC_WAIT_FULL() do { const long long t_ = clock64(); tcw_full += t_ - tcw_a; tcw_a = t_; } while (0)
#define TC_WAIT_CONV() do { const long long t_ = clock64(); tcw_conv += t_ - tcw_a; ++tcw_n; } while (0)
#define TC_WAIT_END() do { if (blockIdx.x == 0 && half == 0) { g_tc_wait[0] = tcw_full; g_tc_wait[1] = tcw_conv; g_tc_wait[2] = clock64() - tcw_t0; g_tc_wait[3] = tcw_n; } } while (0)
extern "C" __attribute__((visibility("default"))) int nvae_debug_tc_wait(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_tc_wait, sizeof(g_tc_wait));
}
extern "C" __attribute__((visibility("default"))) int nvae_debug_tc_cycles(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_tc_cyc, sizeof(g_tc_cyc));
}
extern "C" __attribute__((visibility("default"))) int nvae_debug_tc_timing(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_tc_timing, sizeof(g_tc_timing));
}
#else
#define TC_STAMP(i)
#define TC_CYC(it, j)
#define TC_WAIT_BEGIN()
#define TC_WAIT_A()
#define TC_WAIT_FULL()
#define TC_WAIT_CONV()
#define TC_WAIT_END()
#endif

struct SmemCtl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t conv[kMaxStages];      // lo slot filled by the converters
  uint64_t lo_empty[kMaxStages];  // lo slot consumed by the MMAs
  uint64_t acc_full[2];           // TMEM accumulator double buffer: MMAs of segment i+1 overlap the epilogue of i
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

// Epilogue staging: each epilogue warp transposes its 32 rows x 32 columns through shared memory (16-byte slots,
// XOR-swizzled by row so both the row-per-lane writes and the 8-lanes-per-row reads are conflict-free) so that
// global stores are 128-byte row segments instead of 32 scattered 16-byte pieces.
struct EpiSmem {
  float4 tile[4][32][8];
  long long rowbase[kBM];  // RowCtx.base of each tile row, -1 when the row is not stored
};

__device__ __forceinline__ uint32_t tmem_cols_for(int bn) {
  return bn <= 32 ? 32u : bn <= 64 ? 64u : bn <= 128 ? 128u : 256u;
}
// TMEM columns. TF32: two accumulators of tmem_cols_for(BN).  3xTF32 (BN <= 192): all 512 columns -- accumulators
// packed at [0, BN) and [BN, 2BN), and for forward/dgrad the A operand ring at 384: slot s = {hi 32 cols, lo 32 cols}.
__device__ __forceinline__ uint32_t tmem_alloc_cols(int bn, int passes) {
  return passes == 3 ? 512u : 2u * tmem_cols_for(bn);
}
__device__ __forceinline__ uint32_t tmem_acc_stride(int bn, int passes) {
  return passes == 3 ? (uint32_t)bn : tmem_cols_for(bn);
}
// first TMEM column of the A ring: behind the accumulator(s)
__device__ __forceinline__ uint32_t tmem_a_ring(int bn, int acc_bufs, int dual) { return (uint32_t)((dual ? 2 : acc_bufs) * bn); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// index: operand (0 = A source 1, 1 = A source 2, 2 = B); wgrad only: 3 = x and 4 = dy as 5-D "chunked" views
// {32 ch, W, H, N, C/32} whose box takes several 32-channel chunks in ONE TMA instruction
struct TmapSet {
  CUtensorMap m[5];
};

struct TcParams {
  // pixel tiling
  int N, H, W;
  int tw, th, tn, tiles_h;      // box extent; rows per box = tw*th*tn
  // tiles and the stream-K partition
  int acc_bufs;                 // TMEM accumulators: 2 (a tile's epilogue overlaps the next tile's MMAs) or 1 (long K:
                                // the columns go to a deeper A ring instead, so the converters run further ahead)
  int pair;                     // launched as 2-CTA clusters (cta_group::2): tiles are 256 pixel rows, B is split
  int n_mtiles;                 // 128-row M tiles of the problem (a pair's second tile may lie beyond it)
  int a_tmem;                   // 3xTF32: A (high and low parts) is staged in TMEM by the converters, only B lo in smem
  int BN, stages, lo_stages, passes;  // raw-tile ring, lo-tile ring (3xTF32 only); passes: 1 = TF32, 3 = 3xTF32
  int f16;                      // 3xFP16 (forward / dgrad, a_tmem): both operands are scaled by a power of two that
                                // puts their absmax in [2^14, 2^15), split into fp16 high + low parts by the converters
                                // (A by the converters -> TMEM; B by a pack kernel ahead of the launch, 128-byte rows of
                                // [hi 32 x fp16 | lo 32 x fp16] per 32-deep K chunk that TMA drops into the raw ring) and multiplied
                                // as hi*hi + hi*lo + lo*hi with kind::f16 -- the same 22 significant bits per operand
                                // as 3xTF32 at twice the tensor-pipe rate; the epilogue undoes the scales
  int issue2;                   // two MMA-issuing threads for two-accumulator 3xFP16 tiles (NVAE_TC_ISSUE2=0: one)
  int f16_terms;                // 3: hi*hi + hi*lo + lo*hi.  2 (the >= 20 GFLOP launches unless NVAE_F16X2=0): the A_hi*B_lo term is
                                // not issued, i.e. B enters rounded to fp16 (11 significant bits) -- 2/3 of the MMAs
  const float* amax;            // device: {absmax(A operand), absmax(B operand)}, written just before the launch
  int dual;                     // 3xFP16: a CTA tile is TWO 128-row M tiles (mt = 2*(t / n_ntiles) + g) that share every
                                // staged B tile: stage = [A0][A1][B], converter group g splits A_g, accumulator g at column
                                // g*BN.  Partials / fix-up use the pair layout (p.pair = 1): [cta][slot][g][128][BN]
  int nsub;                     // 3xFP16: the BN-wide tile is nsub accumulators of BN/nsub columns (one MMA each) that share
                                // every staged + converted A tile -- fewer L2 -> shared-memory bytes per MAC
  int n_ntiles;                 // tile t = mt * n_ntiles + nt
  int KU;                       // k-units (pipeline stages) per tile
  int T;                        // tiles
  int whole_tiles;              // 1: CTA ranges are whole tiles (no tile is split); 0: equal unit ranges (stream-K)
  long long U;                  // tiles * KU
  uint32_t a_bytes, b_bytes;    // raw bytes of the A / B part of one stage
  float* part;                  // [gridDim.x][2][128][BN] raw partial accumulators
  // forward / dgrad: K loop and epilogue
  int taps, S;
  int8_t tap_dh[25], tap_dw[25];  // A box origin shift of k-loop tap i (rows / columns of the A pixel grid)
  int8_t tap_par[25];           // stride-2 forward: which (row parity*2 + column parity) plane of x the tap reads
  int8_t tap_w[25];             // weight tap (r*S+s) the k-loop tap i multiplies with
  int x_chunked, dy_chunked;    // wgrad: the four x boxes / the BN/32 dy boxes of a stage come from one 5-D box each
  int a5d;                      // A maps are 5-D stride-2 views {2C, W/2, 2, H/2, N} of x (forward / wgrad of stride 2)
  int par_c;                    // ... channel offset of the odd-column plane (= Cin)
  int oH, oW, os, ooh, oow;     // output pixel of tile pixel (n,h,w): (n, h*os + ooh, w*os + oow) in an oH x oW image
  int nchunk1, nchunk2;         // 32-channel chunks of source 1 / source 2
  int k2_base;                  // K index of source 2's first channel inside one tap (= Cin)
  int bk_tap, br_tap;           // B box origin of tap t: (t*bk_tap + k, t*br_tap + n0)
  int n_valid, n_split;         // columns < n_split -> out1, [n_split, n_valid) -> out2
  float* out1;
  float* out2;
  int ld1, off1, ld2;
  const float* bias;
  const float* res;             // residual laid out like out1
  int accumulate;
  // wgrad
  int KP;                       // pixels per stage (K extent)
  int pad_t, pad_l, njobs;      // job = (tap, 32-channel chunk); 4 jobs per M tile
  int Cin, Cin2, Ct, Cout;
  // operand prolog (1x1 stride-1 forward / backward-filter, 3xTF32): the A operand is act(x * scale[c] + shift[c]),
  // applied by the converter warps on the staged tile -- the activated tensor never exists in memory
  const float* pro;             // scale[pro_C] followed by shift[pro_C] (rows 2, 3 of a BatchNorm stat block), or nullptr
  int pro_C, pro_act;
};

// ---- stream-K partition (identical arithmetic in every role and in the fix-up kernel) -----------------
__device__ __forceinline__ long long cta_u0(const TcParams& p, int c, int G) {
  return p.whole_tiles ? ((long long)c * p.T / G) * p.KU : (long long)c * p.U / G;
}
// first CTA whose range reaches into tile t
__device__ __forceinline__ int first_cta_of(const TcParams& p, int t, int G) {
  const long long x = (long long)t * p.KU;
  int c = (int)(x * G / p.U);
  while (cta_u0(p, c + 1, G) <= x) ++c;
  return c;
}

// Order in which a CTA walks its unit range [u_begin, u_end).  Natural order = ascending units.  For wgrad the K
// axis is the pixel sweep: if the range opens with the tail of a tile (ka > 0) that piece is done LAST, so every
// CTA starts at pixel 0 and all 148 sweep the activations in step -- the working set in L2 is a narrow pixel
// window instead of both whole tensors (which thrash a 126 MB L2: 2 GB of DRAM reads for a 128 MB problem).
struct SegOrder {
  long long lo[2], hi[2];
  int n;
};
__device__ __forceinline__ SegOrder seg_order(long long u_begin, long long u_end, int KU, bool pixel_sweep) {
  SegOrder o;
  const long long first_end = min(u_end, (u_begin / KU + 1) * (long long)KU);
  if (pixel_sweep && u_begin % KU != 0 && first_end < u_end) {
    o.n = 2;
    o.lo[0] = first_end; o.hi[0] = u_end;
    o.lo[1] = u_begin; o.hi[1] = first_end;
  } else {
    o.n = 1;
    o.lo[0] = u_begin; o.hi[0] = u_end;
    o.lo[1] = o.hi[1] = u_end;
  }
  return o;
}

// ---- epilogue addressing ----------------------------------------------------------------------------------
struct RowCtx {
  bool ok;
  int64_t base;  // GEMM: pixel index; WGRAD: weight row (tap*Ct + channel)
};

template <bool WGRAD>
__device__ __forceinline__ RowCtx row_ctx(const TcParams& p, int mt, int row) {
  RowCtx r;
  if (!WGRAD) {
    int n0, h0;
    if (p.tn > 1) { n0 = mt * p.tn; h0 = 0; }
    else { n0 = mt / p.tiles_h; h0 = (mt - n0 * p.tiles_h) * p.th; }
    const int per_img = p.tw * p.th;
    const int in = row / per_img, rem = row - in * per_img;
    const int ih = rem / p.tw, iw = rem - ih * p.tw;
    r.ok = mt < p.n_mtiles && row < per_img * p.tn && (n0 + in) < p.N && (h0 + ih) < p.H;
    r.base = ((int64_t)(n0 + in) * p.oH + (h0 + ih) * p.os + p.ooh) * p.oW + iw * p.os + p.oow;
  } else {
    const int nch = p.nchunk1 + p.nchunk2;
    const int job = mt * 4 + (row >> 5);
    const int tap = job / nch, c = job - tap * nch;
    const bool s2 = c >= p.nchunk1;
    const int ch = (s2 ? c - p.nchunk1 : c) * kChunk + (row & 31);
    r.ok = job < p.njobs && ch < (s2 ? p.Cin2 : p.Cin);
    r.base = (int64_t)tap * p.Ct + (s2 ? p.Cin : 0) + ch;
  }
  return r;
}

// final store of 4 consecutive columns starting at tile column `col` (multiple of 4)
template <bool WGRAD>
__device__ __forceinline__ void store4(const TcParams& p, const RowCtx& r, int nt, int col, float4 o, float oscale) {
  const int n = nt * p.BN + col;
  if (!r.ok || col >= p.BN) return;
  if (p.f16) { o.x *= oscale; o.y *= oscale; o.z *= oscale; o.w *= oscale; }
  if (WGRAD) {
    if (n >= p.Cout) return;
    float* dst = p.out1 + r.base * p.Cout + n;
    if (p.accumulate) {  // a later batch chunk of a chunked backward-filter call
      const float4 b = *reinterpret_cast<const float4*>(dst);
      o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
    }
    stg4(dst, o);
    return;
  }
  if (n >= p.n_valid) return;
  float* dst;
  if (n < p.n_split) {
    dst = p.out1 + r.base * p.ld1 + p.off1 + n;
    if (p.bias) {
      const float4 b = ldg4(p.bias + n);
      o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
    }
    if (p.res) {
      const float4 b = ldg4(p.res + r.base * p.ld1 + p.off1 + n);
      o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
    }
  } else {
    if (p.out2 == nullptr) return;
    dst = p.out2 + r.base * p.ld2 + (n - p.n_split);
  }
  if (p.accumulate) {
    const float4 b = *reinterpret_cast<const float4*>(dst);
    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
  }
  stg4(dst, o);
}

// v - trunc_tf32(v): exact in fp32 (13 significant bits); kind::tf32 reads its top 10, so the dropped part is
// below 2^-21 |v|
__device__ __forceinline__ float tf32_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__device__ __forceinline__ float4 tf32_lo4(const float4& v) {
  return make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
}

// ---- 3xFP16: power-of-two operand scales from the absmax values ------------------------------------------
// absmax = m * 2^e (m in [0.5, 1)) -> scale 2^(15 - e): the scaled absmax lies in [2^14, 2^15), inside fp16's range
// (max 65504) with its low part 2^-11 below still normal down to elements 2^-18 of the absmax
__device__ __forceinline__ int f16_exp(float amax) {
  int e = 0;
  if (amax > 0.f && amax < 3.0e38f) (void)frexpf(amax, &e);
  return e < -100 ? -100 : e;  // (a denormal absmax would ask for a scale beyond fp32's range)
}
__device__ __forceinline__ float f16_in_scale(float amax) { return amax > 0.f ? ldexpf(1.f, 15 - f16_exp(amax)) : 1.f; }
__device__ __forceinline__ float f16_out_scale(const float* amax) {
  const float a = __ldg(amax), b = __ldg(amax + 1);
  return ldexpf(1.f, (a > 0.f ? f16_exp(a) - 15 : 0) + (b > 0.f ? f16_exp(b) - 15 : 0));
}
// 8 consecutive K elements -> packed fp16 high parts (RN) and low parts rn(v - hi)
__device__ __forceinline__ void f16_split8(const float4& v0, const float4& v1, float s, uint4& hi, uint4& lo) {
  const float a[8] = {v0.x * s, v0.y * s, v0.z * s, v0.w * s, v1.x * s, v1.y * s, v1.z * s, v1.w * s};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(a[2 * i], a[2 * i + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(a[2 * i] - hf.x, a[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
// PAIR: forward / dgrad in 3xTF32 as 2-CTA clusters (cta_group::2).  The pair owns a 256-pixel x BN tile: each CTA
// stages its own 128 pixel rows of A (-> its TMEM) and HALF of the B tile (BN/2 weight rows) in its shared memory, the
// leader issues the MMAs for both SMs.  Per SM this halves the B bytes moved by TMA, split by the converters and
// read by the MMAs -- the shared-memory port, not the tensor pipe, is what bounds the single-CTA kernel.
// PRO: the operand prolog (TcParams::pro) is compiled in.  A separate instantiation: even unused, its code in the converter
// loop cost every 3xTF32 launch ~2 % (measured: 29.8 -> 30.5 ms per step), the converters being on the critical path.
template <bool WGRAD, bool PAIR, bool PRO = false>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ TmapSet maps, const TcParams p) {
  nvae::pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem_raw);
  const uint32_t stage_base = (smem_u32(smem_raw) + (uint32_t)sizeof(SmemCtl) + 1023u) & ~1023u;
  const uint32_t raw_bytes = p.a_bytes + p.b_bytes;
  const uint32_t stage_bytes = raw_bytes;                                   // raw ring: [A][B] per slot
  // lo ring.  a_tmem: [B lo] only -- A's high and low parts go to TMEM; else (wgrad with > 32 pixels per stage)
  // [A lo][B lo] per slot
  const uint32_t lo_base = stage_base + (uint32_t)p.stages * raw_bytes;
  const uint32_t lo_bytes = p.f16 ? 0u : (p.a_tmem ? p.b_bytes : raw_bytes);  // 3xFP16: B arrives pre-split, no lo ring
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // PAIR: the work unit space is over 256-row pair tiles and is cut over gridDim.x / 2 pairs
  const int G = PAIR ? gridDim.x >> 1 : gridDim.x, cta = PAIR ? blockIdx.x >> 1 : blockIdx.x;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const long long u_begin = cta_u0(p, cta, G), u_end = cta_u0(p, cta + 1, G);
  const int nch = p.nchunk1 + p.nchunk2;
  const SegOrder so = seg_order(u_begin, u_end, p.KU, WGRAD);

  if (threadIdx.x == 0) {
    TC_STAMP(0);
    // two MMA issuers (3xFP16 tiles with two accumulators): both commit every stage / A slot / accumulator
    const uint32_t nissue = (!PAIR && p.f16 && (p.nsub == 2 || p.dual) && p.issue2) ? 2u : 1u;
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(smem_u32(&ctl->full[i]), 1);
      mbar_init(smem_u32(&ctl->empty[i]), nissue);
    }
    for (int i = 0; i < p.lo_stages; ++i) {
      // a_tmem: one 4-warp converter group per stage (PAIR: of both CTAs); else all 8 converter warps
      mbar_init(smem_u32(&ctl->conv[i]), (PAIR ? 2 : 1) * ((p.a_tmem && !p.dual) ? kConvThreads / 64 : kConvThreads / 32));
      mbar_init(smem_u32(&ctl->lo_empty[i]), nissue);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&ctl->acc_full[i]), nissue);
      mbar_init(smem_u32(&ctl->acc_empty[i]), (PAIR ? 2 : 1) * (kEpiThreads / 32));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_prefetch_desc(&maps.m[0]);
    tma_prefetch_desc(&maps.m[2]);
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_2cta(smem_u32(&ctl->tmem_base), tmem_alloc_cols(p.BN, p.passes));
    else tmem_alloc(smem_u32(&ctl->tmem_base), tmem_alloc_cols(p.BN, p.passes));
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers exist before anything arrives on them remotely
  tc_fence_after();
  nvae::pdl_wait();  // barrier init, TMEM allocation and descriptor prefetch overlap the previous kernel's tail
  const uint32_t tmem = ctl->tmem_base;
  if (threadIdx.x == 0) TC_STAMP(1);

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int it = 0;
      for (int ph = 0; ph < so.n; ++ph)
      for (long long u = so.lo[ph], u_end = so.hi[ph]; u < u_end;) {
        const int t = (int)(u / p.KU), ka = (int)(u - (long long)t * p.KU);
        const int kb = (int)min((long long)p.KU, ka + (u_end - u));
        const bool dual = !PAIR && p.dual;
        const int mt = PAIR ? 2 * (t / p.n_ntiles) + (int)crank : dual ? 2 * (t / p.n_ntiles) : t / p.n_ntiles,
                  nt = t - (t / p.n_ntiles) * p.n_ntiles;
        if (!WGRAD) {
          int n0, h0;
          if (p.tn > 1) { n0 = mt * p.tn; h0 = 0; }
          else { n0 = mt / p.tiles_h; h0 = (mt - n0 * p.tiles_h) * p.th; }
          int n1 = 0, h1 = 0;  // dual: the second M tile (beyond the problem: TMA zero fill)
          if (dual) {
            if (p.tn > 1) { n1 = (mt + 1) * p.tn; h1 = 0; }
            else { n1 = (mt + 1) / p.tiles_h; h1 = (mt + 1 - n1 * p.tiles_h) * p.th; }
          }
          const uint32_t a_tile = (uint32_t)(p.tw * p.th * p.tn) * 128u;
          const uint32_t tx = (dual ? 2u : 1u) * a_tile + p.b_bytes;
          int tap = ka / nch, c = ka - tap * nch;
          for (int k = ka; k < kb; ++k, ++it) {
            const int ah = h0 + p.tap_dh[tap], aw = p.tap_dw[tap], wt = p.tap_w[tap];
            const int st = it % p.stages;
            const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
            mbar_wait(smem_u32(&ctl->empty[st]), ph ^ 1u);
            if (it == 8) TC_STAMP(7);
            const uint32_t full = smem_u32(&ctl->full[st]);
            const uint32_t sa = stage_base + (uint32_t)st * stage_bytes;
            mbar_expect_tx(full, tx);
            int kk;
            if (p.a5d) {
              const int par = p.tap_par[tap];
              tma_load_5d(sa, &maps.m[0], full, (par & 1) * p.par_c + c * kChunk, aw, par >> 1, ah, n0);
              kk = c * kChunk;
            } else if (c < p.nchunk1) {
              tma_load_4d(sa, &maps.m[0], full, c * kChunk, aw, ah, n0);
              if (dual) tma_load_4d(sa + (p.a_bytes >> 1), &maps.m[0], full, c * kChunk, aw, h1 + p.tap_dh[tap], n1);
              kk = c * kChunk;
            } else {
              tma_load_4d(sa, &maps.m[1], full, (c - p.nchunk1) * kChunk, aw, ah, n0);
              kk = p.k2_base + (c - p.nchunk1) * kChunk;
            }
            // PAIR: this CTA's half of the B tile (p.b_bytes covers BN/2 rows)
            tma_load_2d(sa + p.a_bytes, &maps.m[2], full, wt * p.bk_tap + kk,
                        wt * p.br_tap + nt * p.BN + (PAIR ? (int)crank * (p.BN >> 1) : 0));
            if (!PAIR && p.nsub == 2)  // second sub-tile (a TMA box has at most 256 rows)
              tma_load_2d(sa + p.a_bytes + (p.b_bytes >> 1), &maps.m[2], full, wt * p.bk_tap + kk,
                          wt * p.br_tap + nt * p.BN + (p.BN >> 1));
            if (++c == nch) { c = 0; ++tap; }
          }
        } else {
          const uint32_t box_bytes = (uint32_t)p.KP * 128u;
          const int nb = p.BN / kChunk;
          const int jmax = dual ? 8 : 4;
          int njob = p.njobs - mt * 4;
          njob = njob > jmax ? jmax : njob;
          int jc[8], jh[8], jw[8], js[8], jp[8];
          for (int j = 0; j < jmax; ++j) {
            const int job = mt * 4 + j;
            const int tap = job / nch, c = job - tap * nch;
            const int r = tap / p.S, s = tap - r * p.S;
            jh[j] = r - p.pad_t; jw[j] = s - p.pad_l;
            js[j] = c >= p.nchunk1;
            jc[j] = (js[j] ? c - p.nchunk1 : c) * kChunk;
            jp[j] = 0;
            if (p.a5d) {  // input row 2*ho + e: plane e & 1, offset e >> 1 (floor) in the half-resolution grid
              jp[j] = jh[j] & 1;
              jc[j] += (jw[j] & 1) * p.par_c;
              jh[j] >>= 1; jw[j] >>= 1;
            }
          }
          const uint32_t tx = (uint32_t)(njob + nb) * box_bytes;
          // the tile's four jobs are consecutive channel chunks of ONE tap -> a single chunked box fetches them
          const bool x_one = p.x_chunked && njob >= 4 && (mt * 4) / nch == (mt * 4 + 3) / nch;
          const bool x_one2 = dual && p.x_chunked && njob == 8 && (mt * 4 + 4) / nch == (mt * 4 + 7) / nch;
          for (int k = ka; k < kb; ++k, ++it) {
            int n0, h0;
            if (p.tn > 1) { n0 = k * p.tn; h0 = 0; }
            else { n0 = k / p.tiles_h; h0 = (k - n0 * p.tiles_h) * p.th; }
            const int st = it % p.stages;
            const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
            mbar_wait(smem_u32(&ctl->empty[st]), ph ^ 1u);
            const uint32_t full = smem_u32(&ctl->full[st]);
            const uint32_t sa = stage_base + (uint32_t)st * stage_bytes;
            mbar_expect_tx(full, tx);
            // (a wgrad stage used to be ten TMA instructions -- 4 x boxes + 6 dy boxes -- and was bound by their issue
            // cost; the chunked 5-D views fetch each operand with one)
            if (x_one) {
              tma_load_5d(sa, &maps.m[3], full, 0, jw[0], h0 + jh[0], n0, jc[0] / kChunk);
            } else {
              for (int j = 0; j < (njob < 4 ? njob : 4); ++j) {
                if (p.a5d) tma_load_5d(sa + (uint32_t)j * box_bytes, &maps.m[0], full, jc[j], jw[j], jp[j], h0 + jh[j], n0);
                else tma_load_4d(sa + (uint32_t)j * box_bytes, &maps.m[js[j]], full, jc[j], jw[j], h0 + jh[j], n0);
              }
            }
            if (x_one2) {
              tma_load_5d(sa + 4u * box_bytes, &maps.m[3], full, 0, jw[4], h0 + jh[4], n0, jc[4] / kChunk);
            } else {
              for (int j = 4; j < njob; ++j)
                tma_load_4d(sa + (uint32_t)j * box_bytes, &maps.m[js[j]], full, jc[j], jw[j], h0 + jh[j], n0);
            }
            if (p.f16) {
              // packed dY [pixel][Cout/64][hi|lo][64 x fp16]: one box = the tile's BN/64 channel chunks, both parts
              tma_load_5d(sa + p.a_bytes, &maps.m[4], full, 0, 0, h0, n0, nt * (nb >> 1) * 2);
            } else if (p.dy_chunked) {
              tma_load_5d(sa + p.a_bytes, &maps.m[4], full, 0, 0, h0, n0, nt * nb);
            } else {
              for (int b = 0; b < nb; ++b)
                tma_load_4d(sa + p.a_bytes + (uint32_t)b * box_bytes, &maps.m[2], full, nt * p.BN + b * kChunk, 0, h0, n0);
            }
          }
        }
        u += kb - ka;
      }
    }
  } else if (warp == 1 || warp == kIssue2Warp) {
    // ---------------- MMA issuer (PAIR: the leader CTA only) ----------------
    // split: the tile has two accumulators (nsub / dual) and two issuing threads -- warp 1 issues the MMAs of the first,
    // warp 14 those of the second; each waits on the stage's barriers and commits on its own.  One thread issuing all of a
    // stage's MMAs costs it ~75 cycles per MMA plus ~600 cycles of waits, commits and loop overhead (in-kernel timeline:
    // stage period 1 435 cycles against 768 of tensor-pipe time for the eight two-term MMAs).
    const bool split = !PAIR && p.f16 && (p.nsub == 2 || p.dual) && p.issue2;
    const int half = warp == kIssue2Warp ? 1 : 0;
    if (lane == 0 && (!PAIR || crank == 0) && (half == 0 || split)) {
      const uint32_t idesc = WGRAD ? umma_idesc_tf32(p.BN, 1, 1) : umma_idesc_tf32(p.BN, 0, 0);
      // A from TMEM is K-major by construction; a pair MMA spans 256 rows
      const uint32_t idesc_ts = umma_idesc_tf32(p.BN, 0, WGRAD ? 1 : 0, PAIR ? 2 * kBM : kBM);
      const uint32_t box_bytes = (uint32_t)p.KP * 128u;
      const int ksteps = WGRAD ? p.KP / 8 : 4;
      const uint64_t kadv = WGRAD ? 64u : 2u;  // descriptor start-address step per K=8: 8 pixel rows / 32 bytes
      if (p.a_tmem) {
        // 3xTF32 fast path.  This ONE thread paces the tensor pipe (a tcgen05.mma costs it ~70 cycles to issue), so
        // everything else in its loop is kept off the critical path: ring indices, barrier addresses and operand
        // descriptors advance incrementally (no divisions), and the 12 MMAs of a stage are issued back to back.
        constexpr uint64_t kAdv = WGRAD ? 64u : 2u;
        // wgrad 3xFP16: B = packed dY, MN-major fp16, 128B swizzle: atom = 64 channels x 8 pixels (1 KB); the tile's
        // chunks lie [c0 hi][c0 lo][c1 hi]... (4 KB each) -> LBO (next 64 channels) = 8 KB, SBO (next 8 pixels) = 1 KB
        const uint64_t b_desc0 = (WGRAD && p.f16) ? umma_desc_sw128(stage_base + p.a_bytes, 8192, 1024, 2)
                                 : WGRAD          ? umma_desc_sw128(stage_base + p.a_bytes, box_bytes, 512, 1)
                                                  : umma_desc_sw128(stage_base + p.a_bytes, 16, 1024);
        const uint64_t lb_desc0 = WGRAD ? umma_desc_sw128(lo_base, box_bytes, 512, 1) : umma_desc_sw128(lo_base, 16, 1024);
        const uint64_t st_step = (uint64_t)(stage_bytes >> 4), lo_step = (uint64_t)(lo_bytes >> 4);
        const uint32_t a_base = tmem + tmem_a_ring(p.BN, p.acc_bufs, !PAIR && p.dual);
        const uint32_t full0 = smem_u32(&ctl->full[0]), empty0 = smem_u32(&ctl->empty[0]);
        const uint32_t conv0 = smem_u32(&ctl->conv[0]), loe0 = smem_u32(&ctl->lo_empty[0]);
        int st = 0, ls = 0, seg = 0, it = 0;
        uint32_t ph = 0, lph = 0, a_hi = a_base;
        TC_WAIT_BEGIN();
        uint64_t db = b_desc0, lb = lb_desc0;
        (void)it;
        const bool f16 = !PAIR && p.f16;
        const bool three = p.f16_terms != 2;
        const int bns = p.nsub == 2 ? p.BN >> 1 : p.BN;  // columns per MMA / accumulator
        const uint32_t idesc_h = umma_idesc_f16(bns, WGRAD ? 1 : 0);
        const uint64_t sub_step = (uint64_t)(p.b_bytes >> 5);  // second sub-tile's B: half the stage's B bytes, in 16-byte units
        // K = 16 step / low-part offset of the B descriptor (16-byte units): forward/dgrad rows are
        // [hi 2 x 32 B | lo 2 x 32 B]; wgrad steps 16 pixel rows of 128 B and finds the low parts 4 KB further
        const uint64_t hk = WGRAD ? 128u : 2u, hlo = WGRAD ? 256u : 4u;
        const uint32_t a_slot = (f16 && !p.dual) ? 32u : 64u;  // TMEM columns per A ring slot (dual: two tiles x 32)
        for (int sp = 0; sp < so.n; ++sp)
        for (long long u = so.lo[sp], u_end = so.hi[sp]; u < u_end; ++seg) {
          const int t = (int)(u / p.KU), ka = (int)(u - (long long)t * p.KU);
          const int kb = (int)min((long long)p.KU, ka + (u_end - u));
          const int ab = seg % p.acc_bufs;
          const uint32_t acc = tmem + (uint32_t)ab * tmem_acc_stride(p.BN, p.passes);
          mbar_wait(smem_u32(&ctl->acc_empty[ab]), (((uint32_t)(seg / p.acc_bufs)) & 1u) ^ 1u);
          tc_fence_after();
          uint32_t accum = 0;
          for (int k = ka; k < kb; ++k, ++it) {
            TC_CYC(it, 13);
            TC_WAIT_A();
            if (!PAIR) mbar_wait(full0 + 8u * st, ph);  // (PAIR: both CTAs' converters vouch for the tiles)
            TC_CYC(it, 0);
            TC_WAIT_FULL();
            mbar_wait(conv0 + 8u * ls, lph);
            tc_fence_after();
            TC_WAIT_CONV();
            TC_CYC(it, 1);
            if (it == 0) TC_STAMP(2);
            const uint32_t a_lo = a_hi + (f16 ? 16u : 32u);
            if (f16) {
              // pre-split B row (TMA, raw ring): [hi k0..15 | hi k16..31 | lo k0..15 | lo k16..31], 32 bytes each
              TC_CYC(it, 8);
              if (!split || half == 0) {
#pragma unroll
                for (int j = 0; j < 2; ++j) umma_f16_ts(acc, a_hi + 8u * j, db + hk * j, idesc_h, accum | (j > 0));
                TC_CYC(it, 9);
                if (three) {
#pragma unroll
                  for (int j = 0; j < 2; ++j) umma_f16_ts(acc, a_hi + 8u * j, db + hlo + hk * j, idesc_h, 1u);
                }
                TC_CYC(it, 10);
#pragma unroll
                for (int j = 0; j < 2; ++j) umma_f16_ts(acc, a_lo + 8u * j, db + hk * j, idesc_h, 1u);
              }
              if (p.dual && (!split || half == 1)) {  // the second M tile's A (slot columns 32..63) against the same B, into the second accumulator
                const uint32_t acc2 = acc + (uint32_t)p.BN, b_hi = a_hi + 32u, b_lo = a_hi + 48u;
#pragma unroll
                for (int j = 0; j < 2; ++j) umma_f16_ts(acc2, b_hi + 8u * j, db + hk * j, idesc_h, accum | (j > 0));
                if (three) {
#pragma unroll
                  for (int j = 0; j < 2; ++j) umma_f16_ts(acc2, b_hi + 8u * j, db + hlo + hk * j, idesc_h, 1u);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) umma_f16_ts(acc2, b_lo + 8u * j, db + hk * j, idesc_h, 1u);
              }
              if (p.nsub == 2 && (!split || half == 1)) {  // same A slot against the second B sub-tile, into the second accumulator
                const uint32_t acc2 = acc + (uint32_t)bns;
                const uint64_t db2 = db + sub_step;
#pragma unroll
                for (int j = 0; j < 2; ++j) umma_f16_ts(acc2, a_hi + 8u * j, db2 + hk * j, idesc_h, accum | (j > 0));
                if (three) {
#pragma unroll
                  for (int j = 0; j < 2; ++j) umma_f16_ts(acc2, a_hi + 8u * j, db2 + hlo + hk * j, idesc_h, 1u);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) umma_f16_ts(acc2, a_lo + 8u * j, db2 + hk * j, idesc_h, 1u);
              }
              TC_CYC(it, 11);
              umma_commit(loe0 + 8u * ls);
              TC_CYC(it, 2);
              umma_commit(empty0 + 8u * st);
              TC_CYC(it, 12);
            } else if (PAIR) {
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_tf32_ts_2cta(acc, a_hi + 8u * j, db + kAdv * j, idesc_ts, accum | (j > 0));
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_tf32_ts_2cta(acc, a_hi + 8u * j, lb + kAdv * j, idesc_ts, 1u);
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_tf32_ts_2cta(acc, a_lo + 8u * j, db + kAdv * j, idesc_ts, 1u);
              umma_commit_2cta(loe0 + 8u * ls);
              umma_commit_2cta(empty0 + 8u * st);
            } else {
              TC_CYC(it, 8);
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_tf32_ts(acc, a_hi + 8u * j, db + kAdv * j, idesc_ts, accum | (j > 0));
              TC_CYC(it, 9);
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_tf32_ts(acc, a_hi + 8u * j, lb + kAdv * j, idesc_ts, 1u);
              TC_CYC(it, 10);
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_tf32_ts(acc, a_lo + 8u * j, db + kAdv * j, idesc_ts, 1u);
              TC_CYC(it, 11);
              umma_commit(loe0 + 8u * ls);
              TC_CYC(it, 2);
              umma_commit(empty0 + 8u * st);
              TC_CYC(it, 12);
            }
            accum = 1u;
            if (++st == p.stages) { st = 0; ph ^= 1u; db = b_desc0; } else { db += st_step; }
            if (++ls == p.lo_stages) { ls = 0; lph ^= 1u; lb = lb_desc0; a_hi = a_base; } else { lb += lo_step; a_hi += a_slot; }
          }
          if (PAIR) umma_commit_2cta(smem_u32(&ctl->acc_full[ab]));
          else umma_commit(smem_u32(&ctl->acc_full[ab]));
          TC_STAMP(3);
          u += kb - ka;
        }
        TC_WAIT_END();
      } else {
      int it = 0, seg = 0;
      for (int ph = 0; ph < so.n; ++ph)
      for (long long u = so.lo[ph], u_end = so.hi[ph]; u < u_end; ++seg) {
        const int t = (int)(u / p.KU), ka = (int)(u - (long long)t * p.KU);
        const int kb = (int)min((long long)p.KU, ka + (u_end - u));
        const int ab = seg % p.acc_bufs;
        const uint32_t acc = tmem + (uint32_t)ab * tmem_acc_stride(p.BN, p.passes);
        mbar_wait(smem_u32(&ctl->acc_empty[ab]), (((uint32_t)(seg / p.acc_bufs)) & 1u) ^ 1u);
        tc_fence_after();
        for (int k = ka; k < kb; ++k, ++it) {
          const int st = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          TC_CYC(it, 13);
          if (!PAIR) {  // (PAIR: the converters of both CTAs vouch for the tiles through conv[])
            mbar_wait(smem_u32(&ctl->full[st]), ph);
            tc_fence_after();
          }
          TC_CYC(it, 14);
          if (it == 0) TC_STAMP(2);
          const uint32_t sa = stage_base + (uint32_t)st * stage_bytes;
          uint64_t da, db;
          if (WGRAD) {
            // MN-major, 128B swizzle with 32B atoms: 32 channels x 4 pixels per 512-byte atom;
            // LBO = next 32-channel box, SBO = next 4 pixels; one K=8 MMA spans two atoms
            da = umma_desc_sw128(sa, box_bytes, 512, 1);
            db = umma_desc_sw128(sa + p.a_bytes, box_bytes, 512, 1);
          } else {
            da = umma_desc_sw128(sa, 16, 1024);
            db = umma_desc_sw128(sa + p.a_bytes, 16, 1024);
          }
          if (!p.a_tmem) {
            // high parts first: they need nothing from the converters, which work on this stage meanwhile
            for (int j = 0; j < ksteps; ++j)
              umma_tf32(acc, da + kadv * j, db + kadv * j, idesc, (k > ka || j > 0) ? 1u : 0u);
          }
          if (p.passes == 3) {
            const int ls = it % p.lo_stages;
            TC_CYC(it, 0);
            mbar_wait(smem_u32(&ctl->conv[ls]), (uint32_t)(it / p.lo_stages) & 1u);
            tc_fence_after();
            TC_CYC(it, 1);
            if (it == 8) TC_STAMP(14);
            if (it == 9) TC_STAMP(15);
            const uint32_t sl = lo_base + (uint32_t)ls * lo_bytes;
            if (!p.a_tmem) {
              const uint64_t la = umma_desc_sw128(sl, box_bytes, 512, 1);
              const uint64_t lb = umma_desc_sw128(sl + p.a_bytes, box_bytes, 512, 1);
              for (int j = 0; j < ksteps; ++j) umma_tf32(acc, da + kadv * j, lb + kadv * j, idesc, 1u);
              for (int j = 0; j < ksteps; ++j) umma_tf32(acc, la + kadv * j, db + kadv * j, idesc, 1u);
            } else {
              // A (high and low parts) comes from TMEM: only B crosses the shared-memory port, three times
              const uint64_t lb = WGRAD ? umma_desc_sw128(sl, box_bytes, 512, 1) : umma_desc_sw128(sl, 16, 1024);
              const uint32_t a_hi = tmem + (uint32_t)(p.acc_bufs * p.BN) + (uint32_t)ls * 64u, a_lo = a_hi + 32u;
              if (PAIR) {
                for (int j = 0; j < 4; ++j)
                  umma_tf32_ts_2cta(acc, a_hi + 8u * j, db + kadv * j, idesc_ts, (k > ka || j > 0) ? 1u : 0u);
                for (int j = 0; j < 4; ++j) umma_tf32_ts_2cta(acc, a_hi + 8u * j, lb + kadv * j, idesc_ts, 1u);
                for (int j = 0; j < 4; ++j) umma_tf32_ts_2cta(acc, a_lo + 8u * j, db + kadv * j, idesc_ts, 1u);
              } else {
                TC_CYC(it, 8);
                for (int j = 0; j < 4; ++j)
                  umma_tf32_ts(acc, a_hi + 8u * j, db + kadv * j, idesc_ts, (k > ka || j > 0) ? 1u : 0u);
                TC_CYC(it, 9);
                for (int j = 0; j < 4; ++j) umma_tf32_ts(acc, a_hi + 8u * j, lb + kadv * j, idesc_ts, 1u);
                TC_CYC(it, 10);
                for (int j = 0; j < 4; ++j) umma_tf32_ts(acc, a_lo + 8u * j, db + kadv * j, idesc_ts, 1u);
                TC_CYC(it, 11);
              }
            }
            if (PAIR) umma_commit_2cta(smem_u32(&ctl->lo_empty[ls]));
            else umma_commit(smem_u32(&ctl->lo_empty[ls]));
            TC_CYC(it, 2);
          }
          if (PAIR) umma_commit_2cta(smem_u32(&ctl->empty[st]));
          else umma_commit(smem_u32(&ctl->empty[st]));
          TC_CYC(it, 12);
        }
        if (PAIR) umma_commit_2cta(smem_u32(&ctl->acc_full[ab]));
        else umma_commit(smem_u32(&ctl->acc_full[ab]));
        TC_STAMP(3);
        u += kb - ka;
      }
      }  // generic path
    }
  } else if (warp < 2 + kConvThreads / 32) {
    // ---------------- converters (3xTF32): lo = v - trunc_tf32(v) of every staged tile ----------------
    // a_tmem: the two 4-warp groups take alternate stages, so two conversions are in flight -- one group's
    // load -> split -> tcgen05.st / st.shared -> fence -> arrive chain (latency, not issue, bound) overlaps the other's.
    if (p.passes == 3) {
      const int n_units = (int)(u_end - u_begin);
      if (p.a_tmem) {
        const int grp = (warp - 2) >> 2, gt = ((warp - 2) & 3) * 32 + lane;  // group, thread within the group (0..127)
        constexpr int kGT = kConvThreads / 2;
        const int n16b = (int)(p.b_bytes >> 4);
        const bool dual = !PAIR && p.dual;  // both groups work on EVERY stage: group g splits M tile g
        const int it0 = dual ? 0 : grp, itstep = dual ? 1 : 2;
        int st = it0 % p.stages, ls = it0 % p.lo_stages;
        uint32_t ph = (uint32_t)(it0 / p.stages) & 1u, lph = (uint32_t)(it0 / p.lo_stages) & 1u;
        const float f16_sa = p.f16 ? f16_in_scale(__ldg(p.amax)) : 1.f;
        constexpr bool pro = PRO && !PAIR;
        for (int it = it0; it < n_units; it += itstep) {
          if (gt == 0) TC_CYC(it, 3);
          mbar_wait(smem_u32(&ctl->lo_empty[ls]), lph ^ 1u);
          if (gt == 0) TC_CYC(it, 4);
          mbar_wait(smem_u32(&ctl->full[st]), ph);
          if (gt == 0) TC_CYC(it, 5);
          const uint8_t* raw = smem_raw + (stage_base - smem_u32(smem_raw)) + (size_t)st * stage_bytes +
                               (dual ? (size_t)grp * (p.a_bytes >> 1) : 0);
          float4* dst = reinterpret_cast<float4*>(smem_raw + (lo_base - smem_u32(smem_raw)) + (size_t)ls * lo_bytes);
          if (!PAIR && p.f16) {
            // ---- 3xFP16: scale, split into fp16 high / low parts -> TMEM (16 + 16 packed columns); B arrives pre-split ----
            const uint32_t ta = tmem + tmem_a_ring(p.BN, p.acc_bufs, dual) + (uint32_t)ls * (dual ? 64u : 32u) +
                                (dual ? (uint32_t)grp * 32u : 0u) + ((uint32_t)((warp & 3) * 32) << 16);
            {
              uint32_t hi[16], lw[16];
              if (!WGRAD) {
                const int m = (warp & 3) * 32 + lane;
                const float4* arow = reinterpret_cast<const float4*>(raw + m * 128);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  uint4 h, l;
                  f16_split8(arow[(2 * c) ^ (m & 7)], arow[(2 * c + 1) ^ (m & 7)], f16_sa, h, l);
                  hi[4 * c] = h.x; hi[4 * c + 1] = h.y; hi[4 * c + 2] = h.z; hi[4 * c + 3] = h.w;
                  lw[4 * c] = l.x; lw[4 * c + 1] = l.y; lw[4 * c + 2] = l.z; lw[4 * c + 3] = l.w;
                }
              } else {
                // A^T: lane = (job, channel), K = the 32 pixels of the job's box (32-byte-atom swizzle, as in 3xTF32)
                const float* box = reinterpret_cast<const float*>(raw + (size_t)(warp & 3) * p.KP * 128);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                  const int k0 = 2 * c, k1 = 2 * c + 1;
                  const float v0 = box[k0 * 32 + ((((lane >> 3) ^ (k0 & 3)) << 3) | (lane & 7))] * f16_sa;
                  const float v1 = box[k1 * 32 + ((((lane >> 3) ^ (k1 & 3)) << 3) | (lane & 7))] * f16_sa;
                  const __half2 hh = __floats2half2_rn(v0, v1);
                  const float2 hf = __half22float2(hh);
                  const __half2 ll = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                  hi[c] = *reinterpret_cast<const uint32_t*>(&hh);
                  lw[c] = *reinterpret_cast<const uint32_t*>(&ll);
                }
              }
              tmem_st16(ta, hi);
              tmem_st16(ta + 16u, lw);
            }
            tmem_st_wait();
            tc_fence_before();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&ctl->conv[ls]));
            st += itstep; if (st >= p.stages) { st -= p.stages; ph ^= 1u; }
            ls += itstep; if (ls >= p.lo_stages) { ls -= p.lo_stages; lph ^= 1u; }
            continue;
          }
          const uint32_t ta = tmem + (uint32_t)(p.acc_bufs * p.BN) + (uint32_t)ls * 64u + ((uint32_t)((warp & 3) * 32) << 16);
          uint32_t hi[16], lw[16];
          // operand prolog: first channel this thread's values belong to (forward: the unit's K chunk; backward-filter:
          // the (job, lane) channel of the tile, constant over the pixel sweep)
          int pc = 0;
          float psc = 1.f, psh = 0.f;
          if (pro) {
            const long long len0 = so.hi[0] - so.lo[0];
            const long long u = it < len0 ? so.lo[0] + it : so.lo[1] + (it - len0);
            const int t = (int)(u / p.KU);
            if (!WGRAD) {
              pc = (((int)(u - (long long)t * p.KU)) % nch) * kChunk;
            } else {
              pc = ((t / p.n_ntiles * 4 + (warp & 3)) % nch) * kChunk + lane;
              psc = __ldg(p.pro + pc);
              psh = __ldg(p.pro + p.pro_C + pc);
            }
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {  // K columns [16*half, +16) of this thread's TMEM lane
            if (!WGRAD) {
              // A: tile row m (128 B, 8 swizzled 16-byte chunks) -> TMEM lane m
              const int m = (warp & 3) * 32 + lane;
              const float4* arow = reinterpret_cast<const float4*>(raw + m * 128);
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                float4 v = arow[(half * 4 + c) ^ (m & 7)];
                if (pro) {
                  const float4 sc = nvae::ldg4(p.pro + pc + 4 * (half * 4 + c));
                  const float4 sh = nvae::ldg4(p.pro + p.pro_C + pc + 4 * (half * 4 + c));
                  v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
                  v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                  if (p.pro_act == NVAE_ACT_SWISH) {
                    v.x = act_fwd<NVAE_ACT_SWISH>(v.x); v.y = act_fwd<NVAE_ACT_SWISH>(v.y);
                    v.z = act_fwd<NVAE_ACT_SWISH>(v.z); v.w = act_fwd<NVAE_ACT_SWISH>(v.w);
                  } else if (p.pro_act == NVAE_ACT_ELU) {
                    v.x = act_fwd<NVAE_ACT_ELU>(v.x); v.y = act_fwd<NVAE_ACT_ELU>(v.y);
                    v.z = act_fwd<NVAE_ACT_ELU>(v.z); v.w = act_fwd<NVAE_ACT_ELU>(v.w);
                  }
                }
                const float4 l = tf32_lo4(v);
                hi[4 * c] = __float_as_uint(v.x); hi[4 * c + 1] = __float_as_uint(v.y);
                hi[4 * c + 2] = __float_as_uint(v.z); hi[4 * c + 3] = __float_as_uint(v.w);
                lw[4 * c] = __float_as_uint(l.x); lw[4 * c + 1] = __float_as_uint(l.y);
                lw[4 * c + 2] = __float_as_uint(l.z); lw[4 * c + 3] = __float_as_uint(l.w);
              }
            } else {
              // A^T: lane m = (job, channel); K = the 32 pixels of the box.  Box layout: pixel rows of 128 B whose
              // 32-byte units are XOR-swizzled with (row & 3) (128B swizzle, 32-byte atoms)
              const float* box = reinterpret_cast<const float*>(raw + (size_t)(warp & 3) * p.KP * 128);
#pragma unroll
              for (int kk = 0; kk < 16; ++kk) {
                const int k = half * 16 + kk;
                float v = box[k * 32 + ((((lane >> 3) ^ (k & 3)) << 3) | (lane & 7))];
                if (pro) {  // (pixels beyond the batch become act(shift), and meet zero dY rows)
                  v = fmaf(v, psc, psh);
                  if (p.pro_act == NVAE_ACT_SWISH) v = act_fwd<NVAE_ACT_SWISH>(v);
                  else if (p.pro_act == NVAE_ACT_ELU) v = act_fwd<NVAE_ACT_ELU>(v);
                }
                hi[kk] = __float_as_uint(v);
                lw[kk] = __float_as_uint(tf32_lo(v));
              }
            }
            tmem_st16(ta + (uint32_t)half * 16u, hi);
            tmem_st16(ta + 32u + (uint32_t)half * 16u, lw);
          }
          const float4* src = reinterpret_cast<const float4*>(raw + p.a_bytes);
          int i = gt;
          for (; i + 3 * kGT < n16b; i += 4 * kGT) {  // loads first: 4 independent LDS.128 in flight
            const float4 v0 = src[i], v1 = src[i + kGT], v2 = src[i + 2 * kGT], v3 = src[i + 3 * kGT];
            dst[i] = tf32_lo4(v0);
            dst[i + kGT] = tf32_lo4(v1);
            dst[i + 2 * kGT] = tf32_lo4(v2);
            dst[i + 3 * kGT] = tf32_lo4(v3);
          }
          for (; i < n16b; i += kGT) dst[i] = tf32_lo4(src[i]);
          tmem_st_wait();
          tc_fence_before();
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(smem_u32(&ctl->conv[ls]), 0);  // the leader's barrier counts both CTAs
            else mbar_arrive(smem_u32(&ctl->conv[ls]));
          }
          if (gt == 0) TC_CYC(it, 6);
          st += 2; if (st >= p.stages) { st -= p.stages; ph ^= 1u; }
          ls += 2; if (ls >= p.lo_stages) { ls -= p.lo_stages; lph ^= 1u; }
        }
      } else {
        // wgrad with more than 32 pixels per stage: A and B low parts both go to the shared-memory lo ring
        const int ct = (warp - 2) * 32 + lane;
        const int n16 = (int)(raw_bytes >> 4);
        int st = 0, ls = 0;
        uint32_t ph = 0, lph = 0;
        for (int it = 0; it < n_units; ++it) {
          mbar_wait(smem_u32(&ctl->lo_empty[ls]), lph ^ 1u);
          mbar_wait(smem_u32(&ctl->full[st]), ph);
          const float4* src = reinterpret_cast<const float4*>(smem_raw + (stage_base - smem_u32(smem_raw)) + (size_t)st * stage_bytes);
          float4* dst = reinterpret_cast<float4*>(smem_raw + (lo_base - smem_u32(smem_raw)) + (size_t)ls * lo_bytes);
          int i = ct;
          for (; i + 3 * kConvThreads < n16; i += 4 * kConvThreads) {
            const float4 v0 = src[i], v1 = src[i + kConvThreads], v2 = src[i + 2 * kConvThreads],
                         v3 = src[i + 3 * kConvThreads];
            dst[i] = tf32_lo4(v0);
            dst[i + kConvThreads] = tf32_lo4(v1);
            dst[i + 2 * kConvThreads] = tf32_lo4(v2);
            dst[i + 3 * kConvThreads] = tf32_lo4(v3);
          }
          for (; i < n16; i += kConvThreads) dst[i] = tf32_lo4(src[i]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&ctl->conv[ls]));
          if (++st == p.stages) { st = 0; ph ^= 1u; }
          if (++ls == p.lo_stages) { ls = 0; lph ^= 1u; }
        }
      }
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> smem transpose -> global (final tile or raw partial) ----
    const int lg = warp & 3;  // TMEM lane group this warp may read: lanes [32*lg, +32)
    const int row = lg * 32 + lane;
    const float oscale = p.f16 ? f16_out_scale(p.amax) : 1.f;
    EpiSmem* epi = reinterpret_cast<EpiSmem*>(smem_raw + (lo_base - smem_u32(smem_raw)) + (size_t)p.lo_stages * lo_bytes);
    const int q = lane & 7, r0 = lane >> 3;
    int seg = 0;
    for (int ph = 0; ph < so.n; ++ph)
    for (long long u = so.lo[ph], u_end = so.hi[ph]; u < u_end; ++seg) {
      const int t = (int)(u / p.KU), ka = (int)(u - (long long)t * p.KU);
      const int kb = (int)min((long long)p.KU, ka + (u_end - u));
      const bool dual = !PAIR && p.dual;
      const int nt = t - (t / p.n_ntiles) * p.n_ntiles;
      const bool full_tile = ka == 0 && kb == p.KU;
      const int ab = seg % p.acc_bufs;
      mbar_wait(smem_u32(&ctl->acc_full[ab]), ((uint32_t)(seg / p.acc_bufs)) & 1u);
      tc_fence_after();
      if (threadIdx.x == kIssue2Warp * 32 - 1) TC_STAMP(4);
      for (int g = 0; g < (dual ? 2 : 1); ++g) {  // dual: accumulator g = M tile 2*(t / n_ntiles) + g
      const int mt = PAIR ? 2 * (t / p.n_ntiles) + (int)crank : dual ? 2 * (t / p.n_ntiles) + g : t / p.n_ntiles;
      const int hrank = PAIR ? (int)crank : g;
      {
        const RowCtx rc = row_ctx<WGRAD>(p, mt, row);
        epi->rowbase[row] = (rc.ok || !full_tile) ? rc.base : -1;
      }
      // slot 0: continues a tile begun by an earlier CTA; slot 1: begins a tile a later CTA finishes
      // (PAIR / dual: [cta][slot][half][128][BN])
      float* pdst = p.part + ((((int64_t)cta * 2 + (ka > 0 ? 0 : 1)) * ((PAIR || dual) ? 2 : 1) + ((PAIR || dual) ? hrank : 0)) * kBM + lg * 32) * p.BN;
      const uint32_t acc = tmem + (uint32_t)ab * tmem_acc_stride(p.BN, p.passes) + (uint32_t)(g * p.BN) + ((uint32_t)(lg * 32) << 16);
      __syncwarp();
      for (int j = 0; j < p.BN; j += 32) {
        uint32_t v[32];
        tmem_ld32(acc + (uint32_t)j, v);
#pragma unroll
        for (int s4 = 0; s4 < 8; ++s4)
          epi->tile[lg][lane][s4 ^ (lane & 7)] = make_float4(__uint_as_float(v[4 * s4]), __uint_as_float(v[4 * s4 + 1]),
                                                             __uint_as_float(v[4 * s4 + 2]), __uint_as_float(v[4 * s4 + 3]));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 4 * i;
          const float4 o = epi->tile[lg][r][q ^ (r & 7)];
          const int col = j + 4 * q;
          if (full_tile) {
            RowCtx rc;
            rc.base = epi->rowbase[lg * 32 + r];
            rc.ok = rc.base >= 0;
            store4<WGRAD>(p, rc, nt, col, o, oscale);
          } else if (col < p.BN) {
            stg4(pdst + (int64_t)r * p.BN + col, o);
          }
        }
        __syncwarp();
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(smem_u32(&ctl->acc_empty[ab]), 0);
        else mbar_arrive(smem_u32(&ctl->acc_empty[ab]));
      }
      if (threadIdx.x == kIssue2Warp * 32 - 1) TC_STAMP(5);
      u += kb - ka;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // neither CTA may leave (or free TMEM) while the pair's MMAs / arrivals can touch it
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2cta(tmem, tmem_alloc_cols(p.BN, p.passes));
    else tmem_dealloc(tmem, tmem_alloc_cols(p.BN, p.passes));
    if (lane == 0) TC_STAMP(6);
  }
}

// Fix-up: block (c, y) finishes the tile whose K range ENDS inside CTA c's first segment: sums the partials of
// the CTAs that covered it and applies the epilogue.  `rows` tile rows per blockIdx.y.  The partial list of a
// tile (up to 148 deep when one tile is all there is) is cut over KG thread groups -- group g sums partials
// g, g+KG, ... in order, four loads in flight -- and the group sums are combined in group order: a fixed tree,
// so results are bit-repeatable.
constexpr int kFixThreads = 256;
// CTAs whose first segment completes a split tile (host-computed: the fix-up launches one block row per entry
// instead of one per CTA, most of which would exit at once)
struct FixList {
  int n;
  uint8_t cta[148];
};
constexpr int kFixMaxGroups = 8;
template <bool WGRAD>
__global__ void __launch_bounds__(kFixThreads) conv_tc_fixup_kernel(const TcParams p, int G, int rows,
                                                                    const FixList fl) {
  nvae::pdl_enter();
  __shared__ float4 red[kFixThreads];
  const float oscale = p.f16 ? f16_out_scale(p.amax) : 1.f;
  const int c = fl.cta[blockIdx.x];
  const long long u0 = cta_u0(p, c, G), u1 = cta_u0(p, c + 1, G);
  const int t = (int)(u0 / p.KU);
  if (u0 == (long long)t * p.KU) return;               // CTA c starts on a tile boundary
  if (u1 < (long long)(t + 1) * p.KU) return;          // ... or does not finish the tile
  const int cf = first_cta_of(p, t, G);
  // pair launches: c, cf index PAIRS; blockIdx.z picks the CTA of the pair (its 128-row half of the pair tile)
  const int rk = p.pair ? (int)blockIdx.z : 0, nrk = p.pair ? 2 : 1;
  const int mt = p.pair ? 2 * (t / p.n_ntiles) + rk : t / p.n_ntiles, nt = t - (t / p.n_ntiles) * p.n_ntiles;
  if (mt >= p.n_mtiles) return;
  const int bn4 = p.BN / 4, total = rows * bn4, nparts = c - cf + 1;
  const int64_t slot = (int64_t)kBM * p.BN;
  const int64_t blk = (int64_t)blockIdx.y * rows * p.BN;
  // partial j of the tile: j = 0 -> CTA cf's slot 1 (it began the tile), j >= 1 -> CTA cf+j's slot 0
  auto part_ptr = [&](int j) {
    return p.part + (((int64_t)(cf + j) * 2 + (j == 0 ? 1 : 0)) * nrk + rk) * slot + blk;
  };
  int KG = kFixThreads / total;
  KG = KG < 1 ? 1 : (KG > kFixMaxGroups ? kFixMaxGroups : KG);
  if (KG > nparts) KG = nparts;
  if (KG == 1) {
    for (int i0 = threadIdx.x; i0 < total; i0 += 4 * kFixThreads) {
      float4 s[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * kFixThreads;
        s[j] = i < total ? ldg4(part_ptr(0) + (int64_t)i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int k = 1; k < nparts; ++k) {
        const float* pk = part_ptr(k);
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + j * kFixThreads;
          v[j] = i < total ? ldg4(pk + (int64_t)i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[j].x += v[j].x; s[j].y += v[j].y; s[j].z += v[j].z; s[j].w += v[j].w; }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * kFixThreads;
        if (i < total) {
          const int row = blockIdx.y * rows + i / bn4, col = (i % bn4) * 4;
          store4<WGRAD>(p, row_ctx<WGRAD>(p, mt, row), nt, col, s[j], oscale);
        }
      }
    }
    return;
  }
  const int g = threadIdx.x / total, e = threadIdx.x - g * total;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g < KG) {
    int k = g;
    for (; k + 3 * KG < nparts; k += 4 * KG) {
      const float4 v0 = ldg4(part_ptr(k) + (int64_t)e * 4), v1 = ldg4(part_ptr(k + KG) + (int64_t)e * 4),
                   v2 = ldg4(part_ptr(k + 2 * KG) + (int64_t)e * 4), v3 = ldg4(part_ptr(k + 3 * KG) + (int64_t)e * 4);
      s.x += v0.x; s.y += v0.y; s.z += v0.z; s.w += v0.w;
      s.x += v1.x; s.y += v1.y; s.z += v1.z; s.w += v1.w;
      s.x += v2.x; s.y += v2.y; s.z += v2.z; s.w += v2.w;
      s.x += v3.x; s.y += v3.y; s.z += v3.z; s.w += v3.w;
    }
    for (; k < nparts; k += KG) {
      const float4 v = ldg4(part_ptr(k) + (int64_t)e * 4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    red[threadIdx.x] = s;
  }
  __syncthreads();
  if (g == 0) {
    for (int q = 1; q < KG; ++q) {
      const float4 v = red[q * total + e];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const int row = blockIdx.y * rows + e / bn4, col = (e % bn4) * 4;
    store4<WGRAD>(p, row_ctx<WGRAD>(p, mt, row), nt, col, s, oscale);
  }
}

__global__ void round_tf32_kernel(float* __restrict__ p, int64_t n4) {
  nvae::pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = *reinterpret_cast<float4*>(p + 4 * i);
    v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
    stg4(p + 4 * i, v);
  }
}

// absmax of two fp32 tensors in one launch (3xFP16 operand scales): blocks [0, g0) reduce a, the rest b; the result
// slots are zeroed by a memset node ahead of the launch.  max is order-independent, so the atomics stay deterministic
// (bit patterns of non-negative floats are monotonic as unsigned integers).
__global__ void __launch_bounds__(256) absmax2_kernel(const float* __restrict__ a, int64_t na4, const float* __restrict__ b,
                                                      int64_t nb4, int g0, unsigned* __restrict__ out) {
  nvae::pdl_enter();
  const bool second = (int)blockIdx.x >= g0;
  const float* p = second ? b : a;
  const int64_t n4 = second ? nb4 : na4;
  const int64_t nblk = second ? (int64_t)gridDim.x - g0 : g0, blk = second ? (int64_t)blockIdx.x - g0 : blockIdx.x;
  float m = 0.f;
  const int64_t stride = nblk * 256;
  int64_t i = blk * 256 + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    const float4 v0 = ldg4(p + 4 * i), v1 = ldg4(p + 4 * (i + stride)), v2 = ldg4(p + 4 * (i + 2 * stride)),
                 v3 = ldg4(p + 4 * (i + 3 * stride));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fmaxf(fabsf(v0.z), fabsf(v0.w))));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fmaxf(fabsf(v1.z), fabsf(v1.w))));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v2.x), fabsf(v2.y)), fmaxf(fabsf(v2.z), fabsf(v2.w))));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v3.x), fabsf(v3.y)), fmaxf(fabsf(v3.z), fabsf(v3.w))));
  }
  for (; i < n4; i += stride) {
    const float4 v = ldg4(p + 4 * i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  m = warp_max(m);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    m = warp_max(m);
    if (threadIdx.x == 0) atomicMax(out + (second ? 1 : 0), __float_as_uint(m));
  }
}

// 3xFP16 B operand: every 32-float K chunk (128 bytes) of the weight matrix becomes 128 bytes of
// [32 fp16 high parts | 32 fp16 low parts] of the scaled values, so the SAME 2-D tensor map / box / swizzle that
// stages the fp32 matrix stages the split operand.  One thread per 8 K elements.
__global__ void __launch_bounds__(256) f16_pack_b_kernel(const float* __restrict__ w, int64_t n8,
                                                         const float* __restrict__ amax, uint4* __restrict__ out) {
  nvae::pdl_enter();
  const float s = f16_in_scale(__ldg(amax + 1));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = i >> 2;
    const int j = (int)(i & 3);
    uint4 h, l;
    f16_split8(ldg4(w + i * 8), ldg4(w + i * 8 + 4), s, h, l);
    out[q * 8 + j] = h;
    out[q * 8 + 4 + j] = l;
  }
}

// wgrad 3xFP16 B operand: dY [pixels][C] fp32 -> [pixels][C/64][hi|lo][64 x fp16] (same bytes), scaled by the
// absmax slot 1.  One thread per 8 channels.
__global__ void __launch_bounds__(256) f16_pack_dy_kernel(const float* __restrict__ dy, int64_t n8, int C8,
                                                          const float* __restrict__ amax, uint4* __restrict__ out) {
  nvae::pdl_enter();
  const float s = f16_in_scale(__ldg(amax + 1));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / C8;
    const int c8 = (int)(i - pix * C8);  // 8-channel group of the pixel; 8 groups per 64-channel chunk
    uint4 h, l;
    f16_split8(ldg4(dy + i * 8), ldg4(dy + i * 8 + 4), s, h, l);
    uint4* o = out + pix * (int64_t)(2 * C8) + (c8 >> 3) * 16 + (c8 & 7);
    o[0] = h;
    o[8] = l;
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

// widest N tile <= cap that cuts n_total into equal parts (wide MMAs run closest to the tensor-pipe peak).
// cap: 256 in TF32; 192 in 3xTF32, where (128 + BN) x 128 B tiles must fit five times (3 raw + 2 lo slots).
int pick_bn(int n_total, int mult, int cap) {
  const int nt = (n_total + cap - 1) / cap;
  return (int)round_up(ceil_div(n_total, nt), mult);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

// 5-D stride-2 view of an NHWC tensor (H, W even): {2C, W/2, 2, H/2, N}; coordinate 0 = column parity * C + c
int make_map_s2(CUtensorMap* m, const float* base, int N, int H, int W, int C, int tw, int th, int tn,
                CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return NVAE_E_DRIVER;
  cuuint64_t dims[5] = {(cuuint64_t)2 * C, (cuuint64_t)W / 2, 2, (cuuint64_t)H / 2, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)2 * C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)2 * W * C * 4,
                           (cuuint64_t)H * W * C * 4};
  cuuint32_t box[5] = {(cuuint32_t)kChunk, (cuuint32_t)tw, 1, (cuuint32_t)th, (cuuint32_t)tn};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NVAE_OK : NVAE_E_DRIVER;
}

// 5-D chunked view of an NHWC tensor view (C % 32 == 0): {32, W, H, N, C/32}; the box takes `nchunk` 32-channel chunks,
// which land in shared memory chunk-major -- exactly the layout of `nchunk` separate 4-D boxes back to back.
int make_map_chunked(CUtensorMap* m, const float* base, int N, int H, int W, int C, int ld, int tw, int th, int tn,
                     int nchunk, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return NVAE_E_DRIVER;
  cuuint64_t dims[5] = {(cuuint64_t)kChunk, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)(C / kChunk)};
  cuuint64_t strides[4] = {(cuuint64_t)ld * 4, (cuuint64_t)W * ld * 4, (cuuint64_t)H * W * ld * 4, (cuuint64_t)kChunk * 4};
  cuuint32_t box[5] = {(cuuint32_t)kChunk, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn, (cuuint32_t)nchunk};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NVAE_OK : NVAE_E_DRIVER;
}

// wgrad 3xFP16: 5-D view of the packed dY [N,H,W][C/64][hi|lo][64 x fp16] as {64 fp16, W, H, N, 2*C/64}; the box takes
// the `nhalf` (= 2 * BN/64) 128-byte half-chunks of one N tile, which land [c0 hi][c0 lo][c1 hi]... of KP x 128 B each
int make_map_packed_dy(CUtensorMap* m, const void* base, int N, int H, int W, int C, int tw, int th, int tn, int nhalf) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return NVAE_E_DRIVER;
  const cuuint64_t row = (cuuint64_t)C * 4;  // bytes per pixel (hi + lo)
  cuuint64_t dims[5] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)(2 * (C / 64))};
  cuuint64_t strides[4] = {row, (cuuint64_t)W * row, (cuuint64_t)H * W * row, 128};
  cuuint32_t box[5] = {64, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn, (cuuint32_t)nhalf};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NVAE_OK : NVAE_E_DRIVER;
}

// Taps of a stride-2 backward-data launch for the output pixels of parity (a, b): dx[2h'+a, 2w'+b] sums
// dy[h' + (a + pad_t - r)/2, w' + (b + pad_l - s)/2] * w[r, s] over the taps whose offsets are whole.
struct TapList { int n; int8_t dh[25], dw[25], w[25]; };
TapList s2_dgrad_taps(const NvaeConvDesc* d, int a, int b) {
  TapList t{};
  for (int r = 0; r < d->R; ++r)
    for (int s = 0; s < d->S; ++s) {
      const int eh = a + d->pad_t - r, ew = b + d->pad_l - s;
      if ((eh & 1) || (ew & 1)) continue;
      t.dh[t.n] = (int8_t)(eh / 2); t.dw[t.n] = (int8_t)(ew / 2); t.w[t.n] = (int8_t)(r * d->S + s);
      ++t.n;
    }
  return t;
}

// which: 0 forward, 1 dgrad, 2 wgrad
bool common_ok(const NvaeConvDesc* d, int which) {
  if (d->R * d->S > 25) return false;
  if ((d->Cin & 3) || (d->Cin2 & 3) || (d->Cout & 3) || d->Cout < 8) return false;
  if (d->Cin2 > 0 && (d->Cin % kChunk) != 0) return false;
  if ((d->y_ld & 3) || (d->y_off & 3)) return false;
  if (!(d->pre_scale == 0.f && d->pre_shift == 0.f) && !(d->pre_scale == 1.f && d->pre_shift == 0.f)) return false;
  if (d->Wo > 128) return false;
  if (d->stride == 1) return d->Ho == d->H && d->Wo == d->W;
  // stride 2: half-resolution parity planes of x, addressed through a 5-D tensor map
  if (d->stride != 2 || (d->H & 1) || (d->W & 1) || d->Cin2 != 0 || (d->Cin % kChunk) != 0) return false;
  if (d->pad_t < -1 || d->pad_l < -1) return false;
  if (which == 1)  // every output parity class needs a tap (else its pixels would have to be zero-filled)
    for (int ab = 0; ab < 4; ++ab)
      if (s2_dgrad_taps(d, ab >> 1, ab & 1).n == 0) return false;
  return true;
}

// Launch plan shared by the three directions: tiles, stages, stream-K grid, partial-buffer size.
struct Plan {
  PixTile t;
  int BN, stages, lo_stages, a_tmem, acc_bufs, pair, n_mtiles, n_ntiles, KU, G, njobs, f16, nsub, dual;
  size_t pack_bytes;   // 3xFP16: the split B operand (same bytes as the fp32 weights) + 256 for the absmax slots, behind the partials
  uint32_t a_bytes, b_bytes;
  long long U;
  int whole_tiles;
  bool split;          // some tile is shared between CTAs -> partial buffer + fix-up launch
  size_t part_bytes;
  size_t smem;
};

// Work partition.  Splitting a tile's K range over several CTAs costs a partial-accumulator round trip and the
// fix-up launch, so it is done only when the main loop it shortens is longer than that (small spatial scales:
// 18-72 tiles, K up to 2304).  Times in microseconds, calibrated on B200 (one 32-deep K stage ~ 0.95 us in
// 3xTF32, 0.35 us in TF32; fix-up ~ 5 us + partial traffic at ~3 TB/s).
bool finish_plan(Plan* pl, int passes, bool wgrad, bool aligned_k = false) {
  // pair launches: the schedulable tiles are 256-row pair tiles and the workers are the 74 CTA pairs
  const int workers = pl->pair ? kNumSMs / 2 : kNumSMs;
  const long long T = (long long)((pl->pair || pl->dual) ? (pl->n_mtiles + 1) / 2 : pl->n_mtiles) * pl->n_ntiles;
  pl->U = T * pl->KU;
  const double t_unit = passes == 3 ? (pl->f16 ? 0.5 : 0.95) : 0.35;
  const double slot_us = (double)kBM * pl->BN * 4 * 2 / 3.0e6;  // one partial written + read back
  pl->whole_tiles = 1;
  pl->G = (int)(T < workers ? T : workers);
  double best = (double)ceil_div(T, pl->G) * pl->KU * t_unit;
  if (T < workers) {
    for (int S = 2; S * T <= workers && 2 * S <= pl->KU; ++S) {  // S workers per tile, >= 2 stages each
      const double t = (double)ceil_div(pl->KU, S) * t_unit + 5.0 + (double)S * T * slot_us;
      if (t < best) { best = t; pl->whole_tiles = 0; pl->G = (int)(S * T); }
    }
  }
  // (aligned_k -- backward-filter whose x + dY exceed L2, with fewer tiles than SMs -- keeps the S-way split: its ranges
  // start at the same pixels in every tile, so the CTAs sweep x / dY in a few aligned phases that L2 holds; equal unit
  // ranges start anywhere, the working set becomes both whole tensors and every stage goes to HBM: measured 1.13 ms
  // against 0.75 ms for the 5x5 192 -> 192 layer at 32 x 32)
  if (pl->U >= 2 * workers && !(aligned_k && T < workers)) {  // equal unit ranges over all workers (up to two partials each)
    const double sk = (double)ceil_div(pl->U, workers) * t_unit + 5.0 + 2.0 * workers * slot_us;
    if (sk < 0.95 * best) { best = sk; pl->whole_tiles = 0; pl->G = workers; }
  }
  pl->split = !pl->whole_tiles;
  // ---- shared-memory rings and the TMEM split (needs the partition: units per worker) ----
  const size_t raw = (size_t)pl->a_bytes + pl->b_bytes;
  const size_t budget = (size_t)kSmemBudget - 2048;
  size_t lo_slot = 0;
  if (passes == 3) {  // raw ring for the TMA latency + a double-buffered lo ring (wgrad: A and B; fwd/dgrad: B only)
    pl->a_tmem = (!wgrad || pl->a_bytes == (pl->dual ? 8u : 4u) * 32u * 128u) ? 1 : 0;  // wgrad: 32 pixels per stage fit the TMEM A ring
    lo_slot = pl->f16 ? 0 : (pl->a_tmem ? pl->b_bytes : raw);
    pl->lo_stages = 2;
    pl->acc_bufs = 2;
    // Long K segments per CTA: ONE accumulator (the un-overlapped epilogue is noise next to >= 16 stages) and its
    // TMEM columns go to a deeper A ring (BN + lo_stages*64 <= 512), so the converters run several stages ahead of
    // the MMA thread instead of handing it each stage just in time.
    {
      const long long seg_units = pl->KU < pl->U / pl->G ? pl->KU : pl->U / pl->G;
      if (pl->a_tmem && seg_units >= 16) {
        int deep = (512 - pl->BN) / (pl->f16 ? 32 : 64);
        if (deep > (pl->f16 ? 6 : 4)) deep = pl->f16 ? 6 : 4;
        while (deep > 2 && budget < deep * lo_slot + 3 * raw) --deep;
        if (deep > 2) { pl->acc_bufs = 1; pl->lo_stages = deep; }
      }
    }
    if (pl->dual) {  // two accumulators (one per M tile) + A slots of 2 x 32 columns
      pl->acc_bufs = 1;
      pl->lo_stages = (512 - 2 * pl->BN) / 64;
      if (pl->lo_stages > 4) pl->lo_stages = 4;
      if (pl->lo_stages < 2) return false;
    }
    if (pl->nsub == 2) {  // two accumulators side by side fill TMEM up to the A ring: no accumulator double buffer
      pl->acc_bufs = 1;
      pl->lo_stages = (512 - pl->BN) / 32;
      if (pl->lo_stages > 6) pl->lo_stages = 6;
      if (pl->lo_stages < 2) return false;
    }
    if (budget < pl->lo_stages * lo_slot + 2 * raw) return false;
    pl->stages = (int)((budget - pl->lo_stages * lo_slot) / raw);
  } else {
    if (budget < 2 * raw) return false;
    pl->lo_stages = 0;
    pl->acc_bufs = 2;
    pl->a_tmem = 0;
    pl->stages = (int)(budget / raw);
  }
  if (pl->stages > kMaxStages) pl->stages = kMaxStages;
  // The converters may run at most `stages` k-units ahead of the oldest TMA load still in flight: a converter warp is
  // released for unit `it` by lo_empty (MMAs of unit it - lo_stages complete, hence loads <= it - lo_stages landed) and then
  // waits full[it % stages] by PARITY -- which only tells unit `it` from unit it - stages, not from it - 2*stages.  With
  // lo_stages > stages it could pass on the parity of the load before last and split a tile that had not landed yet
  // (found by the bitwise-repeatability test at batch 144: the N = 384 backward-filter ran stages = 3, lo_stages = 4 and
  // ~1 launch in 13 got one 128-row tile wrong by one k-unit).  So the TMEM A ring is never deeper than the raw ring.
  if (passes == 3 && pl->lo_stages > pl->stages) pl->lo_stages = pl->stages;
  pl->smem = sizeof(SmemCtl) + 1024 + (size_t)pl->stages * raw + (size_t)pl->lo_stages * lo_slot + sizeof(EpiSmem);
  pl->part_bytes = pl->split ? al256((size_t)pl->G * 2 * ((pl->pair || pl->dual) ? 2 : 1) * kBM * pl->BN * sizeof(float)) : 0;
  return true;
}

// Measured on B200 (5x5 384->384, batch 144): pairs 1.43 ms vs 1.10 ms for independent CTAs -- each SM's shared memory
// still serves its half of B to BOTH tensor cores, so the MMA-side port traffic does not drop, and the cross-SM
// handshakes lengthen the converter -> MMA chain.  Kept as an opt-in experiment (NVAE_TC_PAIR=1), off by default.
bool pair_enabled() {
  static const bool on = []() {
    const char* e = getenv("NVAE_TC_PAIR");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

// 3xFP16 instead of 3xTF32 (NVAE_PREC_TF32X3 only; same accuracy class, twice the MMA rate, but two extra passes over
// the operands for their absmax): taken by the large GEMMs, where the MMAs are what the time goes to.
// NVAE_F16X3=0 disables it, NVAE_F16X3_MIN_GFLOP moves the threshold (tests set 0 to cover small shapes).
bool use_f16x3(const NvaeConvDesc* d, int which) {
  if (d->precision != NVAE_PREC_TF32X3 || d->stride != 1 || d->Cin2 != 0) return false;
  if (d->y_off != 0 || (d->y_ld != 0 && d->y_ld != d->Cout)) return false;
  if (which == 2) {  // packed dY: 64-channel chunks; x: whole 32-channel jobs
    if (d->Cout % 64 != 0 || d->Cin % kChunk != 0) return false;
    const char* w = getenv("NVAE_F16X3_WGRAD");
    if (w != nullptr && w[0] == '0') return false;
  } else if ((which == 0 ? d->Cin : d->Cout) % kChunk != 0) {
    return false;  // whole 32-deep K chunks (the packed B rows)
  }
  const char* e = getenv("NVAE_F16X3");
  if (e != nullptr && e[0] == '0') return false;
  const char* m = getenv("NVAE_F16X3_MIN_GFLOP");
  const double min_gflop = m != nullptr ? atof(m) : 20.0;
  const double gflop = 2.0 * d->N * d->Ho * d->Wo * (double)d->Cout * d->Cin * d->R * d->S * 1e-9;
  return gflop >= min_gflop;
}

// Two-term product for the LARGE 3xFP16 GEMMs (>= 20 GFLOP per launch: the six 5x5 convolutions of the postprocess tower,
// whatever NVAE_F16X3_MIN_GFLOP says): A_hi*B_hi + A_lo*B_hi, i.e. the B operand (weights; dY for backward-filter) enters
// rounded to nearest fp16 of its amax-scaled value -- 11 significant bits, unbiased.  These GEMMs run at the MMA issue rate
// under the power cap, so 2/3 of the MMAs is the only thing that makes them faster (0.573 -> 0.528 ms).  Adopted on the
// measured whole-step parity at batch 144 (tests/test_step_b144_gpu.py: worst gradient tensor 2.8e-4, median 9.6e-6 against
// the fp32 CUDA-core path; three-term: 5.8e-5 / 8.0e-6; bound 1e-3).  NVAE_F16X2=0 restores the three-term product.
bool f16_two_terms(const NvaeConvDesc* d) {
  const char* e = getenv("NVAE_F16X2");
  if (e != nullptr && e[0] == '0') return false;
  const double gflop = 2.0 * d->N * d->Ho * d->Wo * (double)d->Cout * (d->Cin + d->Cin2) * d->R * d->S * 1e-9;
  return gflop >= 20.0;
}

// which: 0 forward, 1 dgrad; ntaps: K-loop taps (a stride-2 dgrad launch covers one output parity class)
bool plan_gemm(const NvaeConvDesc* d, int which, int ntaps, Plan* pl) {
  if (!common_ok(d, which) || !pick_pix_tile(d->N, d->Ho, d->Wo, kBM, 1, &pl->t)) return false;
  const int passes = d->precision == NVAE_PREC_TF32X3 ? 3 : 1;
  pl->f16 = use_f16x3(d, which) && !pair_enabled();
  const int Ct = d->Cin + d->Cin2;
  const int n_total = which == 0 ? d->Cout : Ct;
  pl->n_mtiles = pl->t.n_tiles;
  // 3xTF32 forward / dgrad run as 2-CTA pairs (N tile a multiple of 32: cta_group::2 MMA granularity)
  pl->pair = passes == 3 && pl->n_mtiles >= 2 && pair_enabled();
  pl->BN = pick_bn(n_total, pl->pair ? 32 : 16, passes == 3 ? 192 : 256);
  pl->nsub = 1;
  {  // 3xFP16, 192 < N <= 384: one tile of two accumulators sharing the A tiles (NVAE_F16X3_NSUB=0 disables)
    const char* e = getenv("NVAE_F16X3_NSUB");
    if (pl->f16 && n_total > 192 && n_total <= 384 && n_total % 32 == 0 && !(e != nullptr && e[0] == '0')) {
      pl->BN = n_total;
      pl->nsub = 2;
    }
  }
  pl->n_ntiles = (int)ceil_div(n_total, pl->BN);
  {  // 3xFP16, N <= 192: two M tiles per CTA share each B tile (NVAE_F16X3_DUAL=0 disables)
    const char* e = getenv("NVAE_F16X3_DUAL");
    pl->dual = pl->f16 && pl->nsub == 1 && pl->BN <= 192 && pl->n_mtiles >= 2 && !(e != nullptr && e[0] == '0');
  }
  const int nch = which == 0 ? (int)(ceil_div(d->Cin, kChunk) + ceil_div(d->Cin2, kChunk)) : (int)ceil_div(d->Cout, kChunk);
  pl->KU = ntaps * nch;
  pl->a_bytes = (pl->dual ? 2 : 1) * kBM * 128;
  pl->b_bytes = (uint32_t)(pl->pair ? pl->BN / 2 : pl->BN) * 128;  // pair: each CTA stages half of the B tile
  pl->njobs = 0;
  if (!finish_plan(pl, passes, false)) return false;
  pl->pack_bytes = 0;
  if (pl->f16) {
    pl->pack_bytes = al256((size_t)d->R * d->S * Ct * d->Cout * sizeof(float)) + 256;
    pl->part_bytes += pl->pack_bytes;
  }
  return true;
}

bool plan_wgrad(const NvaeConvDesc* d, Plan* pl) {
  if (!common_ok(d, 2)) return false;
  pl->pair = 0;
  pl->f16 = use_f16x3(d, 2);
  pl->pack_bytes = 0;
  const int passes = d->precision == NVAE_PREC_TF32X3 ? 3 : 1;
  if (!pick_pix_tile(d->N, d->Ho, d->Wo, 32, 8, &pl->t) && !pick_pix_tile(d->N, d->Ho, d->Wo, 64, 8, &pl->t) &&
      !pick_pix_tile(d->N, d->Ho, d->Wo, 128, 8, &pl->t))
    return false;
  const int KP = pl->t.tw * pl->t.th * pl->t.tn;
  const int nch = (int)(ceil_div(d->Cin, kChunk) + ceil_div(d->Cin2, kChunk));
  pl->njobs = d->R * d->S * nch;
  pl->n_mtiles = (pl->njobs + 3) / 4;
  pl->BN = pick_bn(d->Cout, 32, passes == 3 ? 192 : 256);
  pl->nsub = 1;
  {
    const char* e = getenv("NVAE_F16X3_NSUB");
    if (pl->f16 && KP == 32 && d->Cout > 192 && d->Cout <= 384 && d->Cout % 128 == 0 && !(e != nullptr && e[0] == '0')) {
      pl->BN = d->Cout;
      pl->nsub = 2;
    }
  }
  // at least 2 raw + 2 lo (3xTF32) / 3 raw (TF32) slots in shared memory
  const size_t want = passes == 3 ? 4 : 3;
  while (pl->nsub == 1 && (size_t)(4 + pl->BN / kChunk) * KP * 128 * want > (size_t)kSmemBudget - 2048 && pl->BN > 32) pl->BN -= 32;
  pl->n_ntiles = (int)ceil_div(d->Cout, pl->BN);
  pl->KU = pl->t.n_tiles;
  {
    const char* e = getenv("NVAE_F16X3_DUAL");
    pl->dual = pl->f16 && pl->nsub == 1 && KP == 32 && pl->BN <= 192 && pl->BN % 64 == 0 && d->Cout % pl->BN == 0 &&
               pl->n_mtiles >= 2 && !(e != nullptr && e[0] == '0');
  }
  pl->a_bytes = (pl->dual ? 8u : 4u) * KP * 128u;
  pl->b_bytes = (uint32_t)(pl->BN / kChunk) * KP * 128u;
  // 3xFP16 needs the TMEM A path (32 pixels per stage) and N tiles made of whole 64-channel chunks
  if (pl->f16 && (KP != 32 || pl->BN % 64 != 0 || d->Cout % pl->BN != 0)) pl->f16 = 0;
  const double operand_mb = ((double)d->N * d->H * d->W * d->Cin + (double)d->N * d->Ho * d->Wo * d->Cout) * 4e-6;
  if (!finish_plan(pl, passes, true, pl->f16 && operand_mb > 160.0)) return false;
  if (pl->f16) {
    pl->pack_bytes = al256((size_t)d->N * d->Ho * d->Wo * d->Cout * sizeof(float)) + 256;
    pl->part_bytes += pl->pack_bytes;
  }
  return true;
}

void fill_common(TcParams* p, const NvaeConvDesc* d, const Plan& pl, float* part) {
  p->N = d->N; p->H = d->Ho; p->W = d->Wo;  // the GEMM's pixel grid (== H x W for stride 1)
  p->oH = d->Ho; p->oW = d->Wo; p->os = 1; p->ooh = 0; p->oow = 0;
  p->a5d = 0; p->par_c = d->Cin;
  p->tw = pl.t.tw; p->th = pl.t.th; p->tn = pl.t.tn; p->tiles_h = pl.t.tiles_h;
  p->BN = pl.BN; p->stages = pl.stages; p->lo_stages = pl.lo_stages; p->a_tmem = pl.a_tmem; p->acc_bufs = pl.acc_bufs; p->pair = pl.pair || pl.dual; p->dual = pl.dual; p->n_mtiles = pl.n_mtiles; p->passes = d->precision == NVAE_PREC_TF32X3 ? 3 : 1; p->f16 = pl.f16; p->amax = nullptr; p->nsub = pl.nsub;
  p->f16_terms = (pl.f16 && f16_two_terms(d)) ? 2 : 3;
  {
    const char* e = getenv("NVAE_TC_ISSUE2");
    p->issue2 = !(e != nullptr && e[0] == '0');
  }
  p->n_ntiles = pl.n_ntiles; p->KU = pl.KU; p->U = pl.U;
  p->T = ((pl.pair || pl.dual) ? (pl.n_mtiles + 1) / 2 : pl.n_mtiles) * pl.n_ntiles; p->whole_tiles = pl.whole_tiles;
  p->a_bytes = pl.a_bytes; p->b_bytes = pl.b_bytes;
  p->part = part;
}

// 3xFP16 operand preparation on the launch stream: absmax of both operands (memset node + one launch), then the
// split B operand.  Workspace behind the partials: [packed B][256 B: absmax slots].  Returns the packed B in *bc.
int f16_prepare(TcParams* p, const Plan& pl, const float* a, int64_t na, const float* b, int64_t nb, void* ws,
                cudaStream_t stream, const float** bc) {
  uint8_t* base = reinterpret_cast<uint8_t*>(ws) + pl.part_bytes - pl.pack_bytes;
  unsigned* slot = reinterpret_cast<unsigned*>(base + pl.pack_bytes - 256);
  cudaError_t e = cudaMemsetAsync(slot, 0, 2 * sizeof(unsigned), stream);
  if (e != cudaSuccess) return (int)e;
  auto blocks = [](int64_t n, int per) { int64_t g = ceil_div(n, per); return (int)(g < 1 ? 1 : (g > 4 * kNumSMs ? 4 * kNumSMs : g)); };
  const int g0 = blocks(na / 4, 256 * 8), g1 = blocks(nb / 4, 256 * 8);
  nvae::launch(absmax2_kernel, g0 + g1, 256, 0, stream, a, na / 4, b, nb / 4, g0, slot);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  nvae::launch(f16_pack_b_kernel, blocks(nb / 8, 256 * 2), 256, 0, stream, b, nb / 8, reinterpret_cast<const float*>(slot),
               reinterpret_cast<uint4*>(base));
  NVAE_RETURN_IF_LAUNCH_FAILED();
  p->amax = reinterpret_cast<const float*>(slot);
  *bc = reinterpret_cast<const float*>(base);
  return NVAE_OK;
}

template <bool WGRAD, bool PAIR, bool PRO = false>
int launch_main(const TmapSet& maps, const TcParams& p, const Plan& pl, cudaStream_t stream) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<WGRAD, PAIR, PRO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  if (!PAIR) {
    nvae::launch(conv_tc_kernel<WGRAD, PAIR, PRO>, pl.G, kThreads, pl.smem, stream, maps, p);
    NVAE_RETURN_IF_LAUNCH_FAILED();
    return NVAE_OK;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pl.G);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<WGRAD, PAIR>, maps, p);
  ++nvae_launch_counter;
  return e == cudaSuccess ? NVAE_OK : (int)e;
}

template <bool WGRAD>
int launch(const TmapSet& maps, const TcParams& p, const Plan& pl, cudaStream_t stream) {
  int rc = (!WGRAD && pl.pair) ? launch_main<false, true>(maps, p, pl, stream)
           : p.pro != nullptr  ? launch_main<WGRAD, false, true>(maps, p, pl, stream)
                               : launch_main<WGRAD, false>(maps, p, pl, stream);
  if (rc) return rc;
  if (pl.split) {
    FixList fl;
    fl.n = 0;
    for (int c = 1; c < pl.G; ++c) {  // same arithmetic as cta_u0 with whole_tiles == 0
      const long long u0 = (long long)c * pl.U / pl.G, u1 = (long long)(c + 1) * pl.U / pl.G;
      const long long t = u0 / pl.KU;
      if (u0 != t * pl.KU && u1 >= (t + 1) * pl.KU) fl.cta[fl.n++] = (uint8_t)c;
    }
    if (fl.n > 0) {
      // few split tiles -> fewer rows per block, so the fix-up still fills the machine
      const int nrk = (pl.pair || pl.dual) ? 2 : 1;
      int rows = 32;
      while (rows > 1 && (long long)fl.n * nrk * (kBM / rows) < 2 * kNumSMs) rows >>= 1;
      nvae::launch(conv_tc_fixup_kernel<WGRAD>, dim3(fl.n, kBM / rows, nrk), kFixThreads, 0, stream, p, pl.G, rows, fl);
      NVAE_RETURN_IF_LAUNCH_FAILED();
    }
  }
  return NVAE_OK;
}

}  // namespace

bool nvae_conv_tc_supported(const NvaeConvDesc* d, int which) {
  Plan pl;
  if (which == 2) return plan_wgrad(d, &pl);
  if (which == 1 && d->stride == 2) {
    if (!common_ok(d, 1)) return false;
    for (int ab = 0; ab < 4; ++ab)
      if (!plan_gemm(d, 1, s2_dgrad_taps(d, ab >> 1, ab & 1).n, &pl)) return false;
    return true;
  }
  return plan_gemm(d, which, d->R * d->S, &pl);
}

static int wgrad_chunks(const NvaeConvDesc* d);

bool nvae_conv_tc_plan_info(const NvaeConvDesc* d, int which, int32_t* out) {
  Plan pl;
  bool ok;
  if (which == 2) ok = plan_wgrad(d, &pl);
  else if (which == 1 && d->stride == 2) ok = common_ok(d, 1) && plan_gemm(d, 1, s2_dgrad_taps(d, 0, 0).n, &pl);
  else ok = plan_gemm(d, which, d->R * d->S, &pl);
  if (!ok) return false;
  const int32_t v[16] = {1, pl.BN, pl.n_mtiles, pl.n_ntiles, pl.KU, pl.G, pl.stages, pl.lo_stages, pl.acc_bufs, pl.f16,
                         pl.nsub, pl.dual, pl.split ? 1 : 0, (int32_t)pl.smem, (int32_t)(nvae_conv_tc_ws_bytes(d, which) >> 10),
                         which == 2 ? wgrad_chunks(d) : 1};
  for (int i = 0; i < 16; ++i) out[i] = v[i];
  return true;
}

size_t nvae_conv_tc_ws_bytes(const NvaeConvDesc* d, int which) {
  Plan pl;
  if (which == 2) {
    size_t b = plan_wgrad(d, &pl) ? pl.part_bytes : 0;
    const int nc = wgrad_chunks(d);
    if (nc > 1) {
      NvaeConvDesc s = *d;
      s.N = d->N / nc;
      if (plan_wgrad(&s, &pl) && pl.part_bytes > b) b = pl.part_bytes;
    }
    return b;
  }
  if (which == 1 && d->stride == 2) {
    size_t mx = 0;
    if (!common_ok(d, 1)) return 0;
    for (int ab = 0; ab < 4; ++ab)
      if (plan_gemm(d, 1, s2_dgrad_taps(d, ab >> 1, ab & 1).n, &pl) && pl.part_bytes > mx) mx = pl.part_bytes;
    return mx;
  }
  return plan_gemm(d, which, d->R * d->S, &pl) ? pl.part_bytes : 0;
}

// The operand prolog (A = act(BN(x)) applied by the converter warps) exists for the shapes the cells need it for: 1x1,
// stride 1, one source, whole 32-channel chunks, 3xTF32 with A in TMEM, unpaired CTAs.
bool nvae_conv_tc_prolog_supported(const NvaeConvDesc* d) {
  if (d->precision != NVAE_PREC_TF32X3 || d->R != 1 || d->S != 1 || d->stride != 1 || d->Cin2 != 0 || d->Cin % kChunk != 0)
    return false;
  Plan f, g;
  if (!plan_gemm(d, 0, 1, &f) || !plan_wgrad(d, &g)) return false;
  if (f.f16 || f.pair || f.dual || g.f16 || g.pair || g.dual || !g.a_tmem) return false;
  return wgrad_chunks(d) == 1;
}

int nvae_conv2d_fwd_tc(const NvaeConvDesc* d, const float* x, const float* x2, const float* w_tr, const float* bias,
                       const float* residual, float* y, void* ws, size_t ws_bytes, cudaStream_t stream,
                       const float* pro_stat, int pro_act) {
  Plan pl;
  const int Ct = d->Cin + d->Cin2, taps = d->R * d->S;
  if (!plan_gemm(d, 0, taps, &pl)) return NVAE_E_UNSUPPORTED;
  if (pro_stat && !nvae_conv_tc_prolog_supported(d)) return NVAE_E_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(x2) || !aligned16(w_tr) || !aligned16(bias) || !aligned16(residual) || !aligned16(y) ||
      !aligned16(ws))
    return NVAE_E_UNSUPPORTED;
  if (pl.part_bytes > 0 && (ws == nullptr || ws_bytes < pl.part_bytes)) return NVAE_E_WORKSPACE;
  TcParams p{};
  fill_common(&p, d, pl, reinterpret_cast<float*>(ws));
  p.taps = taps; p.S = d->S;
  for (int r = 0; r < d->R; ++r)
    for (int sx = 0; sx < d->S; ++sx) {
      const int i = r * d->S + sx, eh = r - d->pad_t, ew = sx - d->pad_l;
      p.tap_w[i] = (int8_t)i;
      if (d->stride == 1) {
        p.tap_dh[i] = (int8_t)eh; p.tap_dw[i] = (int8_t)ew; p.tap_par[i] = 0;
      } else {  // input row 2*ho + eh: parity plane eh & 1 at half-resolution offset floor(eh / 2)
        p.tap_dh[i] = (int8_t)(eh >> 1); p.tap_dw[i] = (int8_t)(ew >> 1);
        p.tap_par[i] = (int8_t)(((eh & 1) << 1) | (ew & 1));
      }
    }
  p.a5d = d->stride == 2;
  p.nchunk1 = (d->Cin + kChunk - 1) / kChunk;
  p.nchunk2 = (d->Cin2 + kChunk - 1) / kChunk;
  p.k2_base = d->Cin;
  p.bk_tap = Ct; p.br_tap = 0;
  p.n_valid = d->Cout; p.n_split = d->Cout;
  p.out1 = y; p.out2 = nullptr;
  p.ld1 = d->y_ld > 0 ? d->y_ld : d->Cout; p.off1 = d->y_off; p.ld2 = 0;
  p.bias = bias; p.res = residual; p.accumulate = 0;
  if (pro_stat) { p.pro = pro_stat + 2 * (size_t)d->Cin; p.pro_C = d->Cin; p.pro_act = pro_act; }
  TmapSet maps;
  int rc;
  if (d->stride == 2)
    rc = make_map_s2(&maps.m[0], x, d->N, d->H, d->W, d->Cin, pl.t.tw, pl.t.th, pl.t.tn, CU_TENSOR_MAP_SWIZZLE_128B);
  else
    rc = make_map_nhwc(&maps.m[0], x, d->N, d->H, d->W, d->Cin, d->Cin, pl.t.tw, pl.t.th, pl.t.tn);
  if (rc) return rc;
  if (d->Cin2 > 0) rc = make_map_nhwc(&maps.m[1], x2, d->N, d->H, d->W, d->Cin2, d->Cin2, pl.t.tw, pl.t.th, pl.t.tn);
  else maps.m[1] = maps.m[0];
  if (rc) return rc;
  const float* bsrc = w_tr;
  if (pl.f16) {
    rc = f16_prepare(&p, pl, x, (int64_t)d->N * d->H * d->W * d->Cin, w_tr, (int64_t)d->Cout * taps * Ct, ws, stream, &bsrc);
    if (rc) return rc;
  }
  rc = make_map_2d(&maps.m[2], bsrc, d->Cout, (int64_t)taps * Ct, (pl.pair || pl.nsub == 2) ? pl.BN / 2 : pl.BN);
  if (rc) return rc;
  return launch<false>(maps, p, pl, stream);
}

int nvae_conv2d_dgrad_tc(const NvaeConvDesc* d, const float* dy, const float* w_rnd, float* dx, float* dx2,
                         int accumulate, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (!common_ok(d, 1)) return NVAE_E_UNSUPPORTED;
  if (!aligned16(dy) || !aligned16(w_rnd) || !aligned16(dx) || !aligned16(dx2) || !aligned16(ws))
    return NVAE_E_UNSUPPORTED;
  if (dx == nullptr) return NVAE_E_UNSUPPORTED;
  const int Ct = d->Cin + d->Cin2, taps = d->R * d->S;
  const int ld = d->y_ld > 0 ? d->y_ld : d->Cout;
  const int nclass = d->stride == 2 ? 4 : 1;
  for (int ab = 0; ab < nclass; ++ab) {
    TapList tl{};
    if (d->stride == 2) {
      tl = s2_dgrad_taps(d, ab >> 1, ab & 1);
    } else {  // dx[h,w] = sum dy[h + pad_t - r, w + pad_l - s] * w[r,s]
      tl.n = taps;
      for (int r = 0; r < d->R; ++r)
        for (int sx = 0; sx < d->S; ++sx) {
          const int i = r * d->S + sx;
          tl.dh[i] = (int8_t)(d->pad_t - r); tl.dw[i] = (int8_t)(d->pad_l - sx); tl.w[i] = (int8_t)i;
        }
    }
    Plan pl;
    if (!plan_gemm(d, 1, tl.n, &pl)) return NVAE_E_UNSUPPORTED;
    if (pl.part_bytes > 0 && (ws == nullptr || ws_bytes < pl.part_bytes)) return NVAE_E_WORKSPACE;
    TcParams p{};
    fill_common(&p, d, pl, reinterpret_cast<float*>(ws));
    p.taps = tl.n; p.S = d->S;
    for (int i = 0; i < tl.n; ++i) { p.tap_dh[i] = tl.dh[i]; p.tap_dw[i] = tl.dw[i]; p.tap_w[i] = tl.w[i]; p.tap_par[i] = 0; }
    if (d->stride == 2) { p.oH = d->H; p.oW = d->W; p.os = 2; p.ooh = ab >> 1; p.oow = ab & 1; }
    p.nchunk1 = (d->Cout + kChunk - 1) / kChunk;
    p.nchunk2 = 0;
    p.k2_base = 0;
    p.bk_tap = 0; p.br_tap = Ct;
    p.n_valid = Ct; p.n_split = d->Cin;
    p.out1 = dx; p.out2 = dx2;
    p.ld1 = d->Cin; p.off1 = 0; p.ld2 = d->Cin2;
    p.bias = nullptr; p.res = nullptr; p.accumulate = accumulate;
    TmapSet maps;
    int rc = make_map_nhwc(&maps.m[0], dy + d->y_off, d->N, d->Ho, d->Wo, d->Cout, ld, pl.t.tw, pl.t.th, pl.t.tn);
    if (rc) return rc;
    maps.m[1] = maps.m[0];
    const float* bsrc = w_rnd;
    if (pl.f16) {
      rc = f16_prepare(&p, pl, dy, (int64_t)d->N * d->Ho * d->Wo * d->Cout, w_rnd, (int64_t)taps * Ct * d->Cout, ws, stream, &bsrc);
      if (rc) return rc;
    }
    rc = make_map_2d(&maps.m[2], bsrc, (int64_t)taps * Ct, d->Cout, (pl.pair || pl.nsub == 2) ? pl.BN / 2 : pl.BN);
    if (rc) return rc;
    rc = launch<false>(maps, p, pl, stream);
    if (rc) return rc;
  }
  return NVAE_OK;
}

// The large 3xFP16 filter gradients run on the weight-gradient side stream as persistent 148-CTA kernels that own all of
// shared memory for 0.6-0.75 ms: any convolution the (high-priority) main chain reaches meanwhile cannot start until they
// exit.  Cutting the batch into chunks (dW accumulated across them) bounds that wait to one chunk, so the side stream only
// fills the gaps in which the main stream runs its BatchNorm / SE / depthwise kernels.  Opt-in, NVAE_WGRAD_CHUNKS=n:
// measured 31.1 (4 chunks) / 31.6 (2) / 31.8 (6) vs 31.7 ms/step unchunked -- within run-to-run noise, because the step
// is power-capped: overlapping more work does not raise the clock-limited throughput.  Default 1.
static int wgrad_chunks(const NvaeConvDesc* d) {
  if (!use_f16x3(d, 2) || d->Cin2 != 0 || d->y_off != 0) return 1;
  const char* e = getenv("NVAE_WGRAD_CHUNKS");
  int c = e != nullptr ? atoi(e) : 1;
  if (c < 1) c = 1;
  while (c > 1) {
    NvaeConvDesc s = *d;
    s.N = d->N / c;
    if (d->N % c == 0 && use_f16x3(&s, 2)) break;  // the chunk must itself be a large 3xFP16 problem
    --c;
  }
  return c;
}

static int wgrad_tc_once(const NvaeConvDesc* d, const float* x, const float* x2, const float* dy, float* dw, void* ws,
                         size_t ws_bytes, cudaStream_t stream, int accumulate, const float* pro_stat = nullptr,
                         int pro_act = 0);

int nvae_conv2d_wgrad_tc(const NvaeConvDesc* d, const float* x, const float* x2, const float* dy, float* dw, void* ws,
                         size_t ws_bytes, cudaStream_t stream, const float* pro_stat, int pro_act) {
  const int nc = wgrad_chunks(d);
  if (pro_stat && !nvae_conv_tc_prolog_supported(d)) return NVAE_E_UNSUPPORTED;
  if (nc == 1) return wgrad_tc_once(d, x, x2, dy, dw, ws, ws_bytes, stream, 0, pro_stat, pro_act);
  NvaeConvDesc s = *d;
  s.N = d->N / nc;
  const int ld = d->y_ld > 0 ? d->y_ld : d->Cout;
  for (int c = 0; c < nc; ++c) {
    const int rc = wgrad_tc_once(&s, x + (size_t)c * s.N * d->H * d->W * d->Cin, nullptr,
                                 dy + (size_t)c * s.N * d->Ho * d->Wo * ld, dw, ws, ws_bytes, stream, c > 0);
    if (rc) return rc;
  }
  return NVAE_OK;
}

static int wgrad_tc_once(const NvaeConvDesc* d, const float* x, const float* x2, const float* dy, float* dw, void* ws,
                         size_t ws_bytes, cudaStream_t stream, int accumulate, const float* pro_stat, int pro_act) {
  Plan pl;
  if (!plan_wgrad(d, &pl)) return NVAE_E_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(x2) || !aligned16(dy) || !aligned16(dw) || !aligned16(ws)) return NVAE_E_UNSUPPORTED;
  if (pl.part_bytes > 0 && (ws == nullptr || ws_bytes < pl.part_bytes)) return NVAE_E_WORKSPACE;
  const int Ct = d->Cin + d->Cin2;
  const int ld = d->y_ld > 0 ? d->y_ld : d->Cout;
  TcParams p{};
  fill_common(&p, d, pl, reinterpret_cast<float*>(ws));
  p.KP = pl.t.tw * pl.t.th * pl.t.tn;
  p.S = d->S; p.pad_t = d->pad_t; p.pad_l = d->pad_l;
  p.nchunk1 = (d->Cin + kChunk - 1) / kChunk;
  p.nchunk2 = (d->Cin2 + kChunk - 1) / kChunk;
  p.njobs = pl.njobs;
  p.Cin = d->Cin; p.Cin2 = d->Cin2; p.Ct = Ct; p.Cout = d->Cout;
  p.out1 = dw;
  p.accumulate = accumulate;
  p.a5d = d->stride == 2;
  if (pro_stat) { p.pro = pro_stat + 2 * (size_t)d->Cin; p.pro_C = d->Cin; p.pro_act = pro_act; }
  TmapSet maps;
  const CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  int rc;
  if (d->stride == 2) rc = make_map_s2(&maps.m[0], x, d->N, d->H, d->W, d->Cin, p.tw, p.th, p.tn, swz);
  else rc = make_map_nhwc(&maps.m[0], x, d->N, d->H, d->W, d->Cin, d->Cin, p.tw, p.th, p.tn, swz);
  if (rc) return rc;
  if (d->Cin2 > 0) rc = make_map_nhwc(&maps.m[1], x2, d->N, d->H, d->W, d->Cin2, d->Cin2, p.tw, p.th, p.tn, swz);
  else maps.m[1] = maps.m[0];
  if (rc) return rc;
  rc = make_map_nhwc(&maps.m[2], dy + d->y_off, d->N, d->Ho, d->Wo, d->Cout, ld, p.tw, p.th, p.tn, swz);
  if (rc) return rc;
  maps.m[3] = maps.m[0];
  maps.m[4] = maps.m[2];
  // one TMA instruction per operand and stage where the channel counts allow the chunked view
  p.x_chunked = d->stride == 1 && d->Cin2 == 0 && d->Cin % kChunk == 0 && d->Cin >= 4 * kChunk;
  p.dy_chunked = d->Cout % kChunk == 0 && (ld % 4) == 0 && pl.BN % kChunk == 0 && d->Cout % pl.BN == 0;
  if (p.x_chunked) {
    rc = make_map_chunked(&maps.m[3], x, d->N, d->H, d->W, d->Cin, d->Cin, p.tw, p.th, p.tn, 4, swz);
    if (rc) return rc;
  }
  if (p.dy_chunked) {
    rc = make_map_chunked(&maps.m[4], dy + d->y_off, d->N, d->Ho, d->Wo, d->Cout, ld, p.tw, p.th, p.tn, pl.BN / kChunk, swz);
    if (rc) return rc;
  }
  if (pl.f16) {
    // absmax(x), absmax(dY) -> scales; dY is split + re-laid out once (one extra pass over it), x by the converters
    uint8_t* base = reinterpret_cast<uint8_t*>(ws) + pl.part_bytes - pl.pack_bytes;
    unsigned* slot = reinterpret_cast<unsigned*>(base + pl.pack_bytes - 256);
    cudaError_t e = cudaMemsetAsync(slot, 0, 2 * sizeof(unsigned), stream);
    if (e != cudaSuccess) return (int)e;
    auto blocks = [](int64_t n, int per) { int64_t g = ceil_div(n, per); return (int)(g < 1 ? 1 : (g > 4 * kNumSMs ? 4 * kNumSMs : g)); };
    const int64_t nx = (int64_t)d->N * d->H * d->W * d->Cin, ndy = (int64_t)d->N * d->Ho * d->Wo * d->Cout;
    const int g0 = blocks(nx / 4, 256 * 8), g1 = blocks(ndy / 4, 256 * 8);
    nvae::launch(absmax2_kernel, g0 + g1, 256, 0, stream, x, nx / 4, dy, ndy / 4, g0, slot);
    NVAE_RETURN_IF_LAUNCH_FAILED();
    nvae::launch(f16_pack_dy_kernel, blocks(ndy / 8, 256 * 2), 256, 0, stream, dy, ndy / 8, d->Cout / 8,
                 reinterpret_cast<const float*>(slot), reinterpret_cast<uint4*>(base));
    NVAE_RETURN_IF_LAUNCH_FAILED();
    p.amax = reinterpret_cast<const float*>(slot);
    rc = make_map_packed_dy(&maps.m[4], base, d->N, d->Ho, d->Wo, d->Cout, p.tw, p.th, p.tn, 2 * (pl.BN / 64));
    if (rc) return rc;
  }
  return launch<true>(maps, p, pl, stream);
}

int nvae_round_tf32_inplace(float* p, int64_t n, cudaStream_t stream) {
  if (n <= 0) return NVAE_OK;
  if ((n & 3) || !aligned16(p)) return NVAE_E_BADSHAPE;
  int64_t g = ceil_div(n / 4, 256);
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  nvae::launch(round_tf32_kernel, (int)g, 256, 0, stream, p, n / 4);
  NVAE_RETURN_IF_LAUNCH_FAILED();
  return NVAE_OK;
}
