// Sparse 3-D convolution forward / dgrad as an implicit GEMM whose operands are staged by TMA:
//   A (gathered feature rows)  : cp.async.bulk.tensor.2d ... tile::gather4  (UTMALDG.2D.GATHER4) —
//                                one instruction fetches the 32-channel slice of FOUR rulebook rows
//                                into four 128-byte rows of the SWIZZLE_128B operand tile; a missing
//                                neighbour is an out-of-bounds row index, which TMA zero-fills
//                                without touching memory
//   B (filter slice Wt[k])     : one tiled TMA box [cout rows x 32 channels] per (offset, chunk)
//   D                          : tcgen05.mma kind::tf32, fp32 accumulators in TMEM
//
// Why: the cp.async build (sparse_conv_tc.cu) issues 1024 16-byte LDGSTS per 16 KB stage from 128
// producer threads; ncu shows it neither L2- nor tensor-bound but limited by that instruction stream
// (profiles/r1_ncu_spconv_msda_cpasync.md).  Here a stage is 32 gather4 instructions from one warp.
//
// A CTA owns T (1..4) consecutive 128-row output tiles, one TMEM accumulator each, and walks
// (kernel offset k, 32-channel chunk c) in the OUTER loop: the filter slice B(k, c) is fetched once
// and reused by the T tiles, which divides the filter traffic (as large as the gather traffic for
// 64/128-channel layers at 128 rows per CTA) by T.  Offsets no row of a tile uses are skipped.
//
// A-stage ring: the 8 slots are divided into one ring per tile (8 / T slots each).  The issuer of a
// tile is the only consumer of its ring and walks it in order, so the usual "at most one phase
// apart" parity argument holds per ring.  (One ring shared by all tiles does not work: an issuer
// whose tile skips several offsets runs two ring phases ahead of the producers and a parity wait
// cannot tell phase n from phase n + 2.)
//
// Warp roles (416 threads).  UTMALDG takes its operands from uniform registers, so per-lane gather4s
// are serialised inside a warp (about 60 cycles each); the measured issue rate only reaches the L2
// limit (about 14 cycles per 512-byte gather4 per SM, tools/probes/tma_gather4_bench.cu) with 16
// warps issuing concurrently.  Hence:
//   warps 0..7  : rulebook-table loader, then A producers (cp.async: 4 x 16 bytes per thread and
//                 stage; gather4: 2 groups of 4 warps, 8 lanes x 1 gather4 per warp), then epilogue
//                 (tcgen05.ld -> +bias -> global; TMEM lane quadrant = warp & 3, the (tile, 16-column)
//                 items of a quadrant are split over its 2 warps)
//   warps 8..11 : MMA issuers, one per tile of the CTA (own accumulator; one elected lane each).  A
//                 single issuing thread needs about 1200 cycles per stage (tcgen05.mma operands go
//                 through uniform registers: ELECT + R2UR per instruction), which capped the first
//                 build; four issuers run in parallel.  Warp 8 also owns the TMEM allocation.
//   warp 12     : B producer (one tiled TMA per filter slice)
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

#ifdef DDF_PHASES
// instrumented builds only (tools/build_trace.py -DDDF_PHASES): per-CTA phase clocks of the conv kernel
__device__ unsigned long long g_ph[16384][10];
__device__ unsigned int g_ph_n;
extern "C" int ddf_phase_dump(const char* path) {
  static unsigned long long h[16384][10];
  unsigned int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, g_ph_n, sizeof(n));
  cudaMemcpyFromSymbol(h, g_ph, sizeof(h));
  FILE* f = fopen(path, "w");
  if (!f) return 1;
  if (n > 16384) n = 16384;
  for (unsigned i = 0; i < n; ++i) {
    for (int j = 0; j < 10; ++j) fprintf(f, "%llu%c", h[i][j], j == 9 ? '\n' : ',');
  }
  fclose(f);
  n = 0;
  cudaMemcpyToSymbol(g_ph_n, &n, sizeof(n));
  return 0;
}
#endif

namespace {
constexpr int TM = 128;            // rows per tile = UMMA M
constexpr int KCH = 32;            // floats per K chunk (one 128-byte swizzle row)
constexpr int kMaxT = 4;           // tiles per CTA
constexpr int kMaxKvol = 27;
constexpr int kProducerWarps = 8;
constexpr int kGroups = 2;         // gather4 producer groups; group g fills A stages n with n % 2 == g
constexpr int kThreads = (kProducerWarps + kMaxT + 1) * 32;   // + one MMA issuer warp per tile + B producer
constexpr int kStagesA = 8;         // at most; split into per-tile rings of n_slots / T slots (see below)
constexpr int kMaxStagesB = 4;
constexpr int kABytes = TM * KCH * 4;  // 16 KB
constexpr int kDefaultRowsTma = 0;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Bounded wait: a transaction that never completes (a broken tensor map, say) becomes a trap the
// host sees as a launch error instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 24)) {
      printf("spconv mbarrier timeout: block %d thread %d smem 0x%x parity %u\n", blockIdx.x, threadIdx.x, addr, parity);
      __trap();
    }
  } while (!done);
}
// Whole-warp wait.  Every lane polls: electing one lane to poll and parking the others at a warp barrier was
// measured 1.5-2x SLOWER on every conv kernel (B200, tools/bench_ops.py spconv; DDF_POLL_ONE keeps that variant);
// a suspend-time hint on try_wait and nanosleep back-off of the non-critical waiters changed nothing
// (profiles/r2_conv_analysis.md).
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
#ifdef DDF_POLL_ONE
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
#else
  mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int col,
                                            int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ void tma_tile_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int col,
                                            int row) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(smem_u32(bar)), "r"(col), "r"(row)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-convergent variants: every lane executes the statement, one elected lane issues.  With
// operands the compiler can prove warp-uniform they stay in uniform registers, which avoids the
// ELECT / R2UR / branch sequence it otherwise wraps around every UTCHMMA / UTCBAR of a
// single-thread issuer (about 25 SASS instructions per MMA).
__device__ __forceinline__ void umma_tf32_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 with bf16 operands (fp32 accumulate), K = 16 per instruction
__device__ __forceinline__ void umma_bf16_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(
          smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ constexpr uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

// c_format f32, a/b_format bf16 (mma_sm100 instruction descriptor: bits [4,6) D, [7,10) A, [10,13) B)
__device__ __forceinline__ constexpr uint32_t make_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

template <int CO>
struct Cfg {
  // deep enough that a filter slice is requested several TMA latencies (about 2.5 us) ahead of its use
  static constexpr int kStagesB = CO >= 128 ? 2 : 4;   // large configuration (1 CTA per SM)
  static constexpr int kBBytes = CO * KCH * 4;
  static constexpr int kCols = CO < 32 ? 32 : CO;     // TMEM columns per tile
  static constexpr int kIdxBytes = kMaxT * TM * kMaxKvol * 4;
  static constexpr int kSmemBytes = kStagesA * kABytes + kStagesB * kBBytes + kIdxBytes + 512 + 1024;
};

// GATHER4 = false: the A tile is gathered by cp.async (LDGSTS, 16 bytes per thread, zero-fill for
// missing neighbours) from all 16 producer warps; true: by TMA gather4 (kept selectable: measured
// 1.7x slower than cp.async here because UTMALDG issue, not bandwidth, limits it).
//
// PREC = 0: operands are fp32 rows, MMA kind::tf32 (the hardware reads the top 19 bits).
// PREC = 1 ("bf16x3"): every 128-byte chunk of an operand row holds 32 channels as [32 x bf16 hi | 32 x bf16 lo]
// with hi = bf16(x), lo = bf16(x - hi) (ddf_split_bf16x3; same bytes per row as fp32, so the gather, the TMA
// boxes and the stage bookkeeping are unchanged).  Per stage the issuer runs hi.hi + hi.lo + lo.hi as six
// kind::f16 MMAs (K = 16): products carry a 16-bit significand (error about 2^-17 per product instead of 2^-11
// for tf32) at 1.5x the tensor-pipe time of the tf32 stage, fp32 accumulation in TMEM either way.
template <int CO, bool GATHER4, int PREC, int RT>
__global__ void __launch_bounds__(kThreads)
spconv_tma_kernel(const __grid_constant__ CUtensorMap map_feat, const __grid_constant__ CUtensorMap map_w,
                  const float* __restrict__ feat, const int* __restrict__ table, const float* __restrict__ bias, float* __restrict__ out,
                  int n_out, int n_in, int kvol, int cin, int cout, int T, int n_slots, int n_sb) {
  ddf::pdl_sync();
  using C = Cfg<CO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_base = smem;
  uint8_t* b_base = smem + n_slots * kABytes;
  int* s_idx = reinterpret_cast<int*>(b_base + n_sb * C::kBBytes);   // [kvol][T*TM]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_idx) + T * TM * kMaxKvol * 4);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kStagesA;
  uint64_t* b_full = a_empty + kStagesA;
  uint64_t* b_empty = b_full + kMaxStagesB;
  uint64_t* accum_bar = b_empty + kMaxStagesB;
  uint64_t* tbl_bar = accum_bar + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(tbl_bar + 1);
  uint32_t* s_mask = s_tmem + 1;   // [kMaxT]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef DDF_TRACE
  __shared__ long long s_tr[11][64];
  int tr_b = 0;
#define TR(ev, i) do { if (blockIdx.x == 3 && (i) < 64) s_tr[ev][i] = clock64(); } while (0)
#else
#define TR(ev, i) do {} while (0)
#endif
#ifdef DDF_PHASES
  unsigned long long ph[8];
#define PH(i) do { if (tid == 0) ph[i] = clock64(); } while (0)
#else
#define PH(i) do {} while (0)
#endif
  PH(0);
  const int tile0 = blockIdx.x * T;
  const int rows = T * TM;         // rows of this CTA (padded)
  const int vrows = (int)min((long long)rows, (long long)n_out - (long long)tile0 * TM);   // rows that exist
  const int* tbl_src = table + (long long)tile0 * TM * kvol;
  const uint32_t tbl_bulk = ((uint32_t)vrows * (uint32_t)kvol * 4u) & ~15u;   // bytes the bulk copy brings
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(T * C::kCols)) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < kStagesA; ++s) {
      // gather4: one arrive.expect_tx per warp of the owning group; cp.async: one arrival per thread
      // of the owning 2-warp group
      mbar_init(a_full + s, GATHER4 ? kProducerWarps / kGroups : 64 + (RT > 0 ? 1 : 0));
      mbar_init(a_empty + s, 1);
    }
    for (int s = 0; s < n_sb; ++s) {
      mbar_init(b_full + s, 1);
      mbar_init(b_empty + s, T);   // every issuer releases every filter slice
    }
    mbar_init(accum_bar, T);
    mbar_init(tbl_bar, 1);
    for (int t = 0; t < kMaxT; ++t) s_mask[t] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // rulebook rows of the T tiles: the slice [vrows x kvol] of the table is contiguous in global memory -> ONE bulk
    // copy (16-byte granules) into the still unused A-slot area; the producers transpose it into s_idx from there.
    // The strided per-thread global table walk this replaces was a fifth of a CTA's life.
    if (tbl_bulk) {
      mbar_expect_tx(tbl_bar, tbl_bulk);
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(a_base)),
          "l"(tbl_src), "r"(tbl_bulk), "r"(smem_u32(tbl_bar))
          : "memory");
    } else {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(tbl_bar)) : "memory");
    }
  }
  if (warp == kProducerWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  PH(1);
  if (warp < kProducerWarps) {
    // staged table slice [row][k] -> s_idx [k][row] (what the producers read with 16-byte loads); per-tile mask of
    // the offsets in use.  Row stride kvol = 27 is odd: the staged reads are conflict-free.
    mbar_wait(tbl_bar, 0);
    const int* staged = reinterpret_cast<const int*>(a_base);
    for (int r = tid; r < rows; r += kProducerWarps * 32) {
      uint32_t m = 0;
      for (int k = 0; k < kvol; ++k) {
        const uint32_t e = (uint32_t)(r * kvol + k);
        // the at most 3 trailing ints of a ragged last CTA come straight from global memory
        const int j = r < vrows ? (e * 4u < tbl_bulk ? staged[e] : __ldg(tbl_src + e)) : -1;
        s_idx[k * rows + r] = j >= 0 ? j : n_in;   // n_in = out of bounds = zero row
        if (j >= 0) m |= 1u << k;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, off);
      if (lane == 0 && m) atomicOr(s_mask + (r >> 7), m);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  PH(2);
  const uint32_t tmem_base = *s_tmem;
  uint32_t tmask[kMaxT];
  uint32_t any = 0;
#pragma unroll
  for (int t = 0; t < kMaxT; ++t) {
    tmask[t] = t < T ? s_mask[t] : 0u;
    any |= tmask[t];
  }
  const int n_chunks = cin / KCH;
  // A slots per tile ring: n_slots / T rounded down to a power of two (8 slots: 8, 4, 2, 2 for T = 1..4)
  const int ring_shift = (n_slots == 8 ? 3 : 2) - (T == 1 ? 0 : T == 2 ? 1 : 2);
  const int ring = 1 << ring_shift;

  if (warp < kProducerWarps) {
    // ===================== A producers =====================
    int n = 0;                                   // running A-stage number
    int cnt[kMaxT] = {0, 0, 0, 0};               // stages issued so far per tile ring
    if constexpr (GATHER4) {
      const int grp = warp >> 2, wq = warp & 3;  // group, 32-row quarter of the tile
      for (int k = 0; k < kvol; ++k) {
        if (!((any >> k) & 1u)) continue;
        for (int c = 0; c < n_chunks; ++c) {
#pragma unroll
          for (int t = 0; t < kMaxT; ++t) {
            if (!((tmask[t] >> k) & 1u)) continue;
            const int slot = t * ring + (cnt[t] & (ring - 1));
            const uint32_t par = (uint32_t)(cnt[t] >> ring_shift) & 1u;
            ++cnt[t];
            if ((n & (kGroups - 1)) == grp) {
              int4 idx = make_int4(0, 0, 0, 0);
              if (lane < 8) idx = *reinterpret_cast<const int4*>(s_idx + k * rows + t * TM + wq * 32 + lane * 4);
              if (lane == 0) {
                mbar_wait(a_empty + slot, par ^ 1u);
                mbar_expect_tx(a_full + slot, kABytes / 4);
              }
              __syncwarp();
              if (lane < 8)
                tma_gather4(smem_u32(a_base + slot * kABytes) + (wq * 32 + lane * 4) * 128, &map_feat,
                            a_full + slot, c * KCH, idx.x, idx.y, idx.z, idx.w);
            }
            ++n;
          }
        }
      }
    } else {
      // Producer groups of 2 warps; group g owns the A slots s with s % 4 == g, so four stages are
      // being issued at any time (a thread spends about 500 cycles per stage between the barrier
      // poll and 16 dependent address computations; with every thread on every stage that latency
      // was the stage period).  Thread -> 16-byte chunk ch of 16 consecutive rows.
      // Hybrid gather (RT > 0, DDF_CONV_ROWS_TMA=32|64; off by default): the last RT rows of every stage go through
      // the TMA unit (RT / 4 gather4 instructions per stage) in parallel with the LDGSTS of the first TM - RT rows;
      // both land on the same a_full barrier (64 cp.async arrivals + one expect_tx arrival).  Measured: no gain
      // (32 rows: +-1%, 64 rows: 3-15% slower) - the kernels are bound by the gathered bytes themselves, not by the
      // unit that moves them (profiles/r2_conv_analysis.md).
      constexpr int kCpGroups = kProducerWarps / 2;
      constexpr int NR = (TM - RT) / 8;                // LDGSTS rows per thread: 16, 12 or 8
      constexpr int LPW = RT / 8;                      // lanes per warp that issue one gather4 each
      const int grp = warp >> 1, gt = tid & 63;
      const int rb = (gt >> 3) * NR, ch = gt & 7;      // first row, chunk
      const uint32_t a0 = smem_u32(a_base);
      const float* col0 = feat + ch * 4;
      for (int k = 0; k < kvol; ++k) {
        if (!((any >> k) & 1u)) continue;
        const int* ik = s_idx + k * rows;
        for (int c = 0; c < n_chunks; ++c) {
          const float* col = col0 + c * KCH;
#pragma unroll
          for (int t = 0; t < kMaxT; ++t) {
            if (!((tmask[t] >> k) & 1u)) continue;
            const int slot = t * ring + (cnt[t] & (ring - 1));
            const uint32_t par = (uint32_t)(cnt[t] >> ring_shift) & 1u;
            ++cnt[t];
            // a slot always belongs to the same group: one producer and one consumer per slot, both
            // in order, so a parity wait can never be two phases off
            if ((slot & (kCpGroups - 1)) != grp) continue;
            int4 j4[NR / 4];
#pragma unroll
            for (int i = 0; i < NR / 4; ++i) j4[i] = *reinterpret_cast<const int4*>(ik + t * TM + rb + 4 * i);
            const int* j = reinterpret_cast<const int*>(j4);
            int4 jt = make_int4(0, 0, 0, 0);
            int rt0 = 0;
            if constexpr (RT > 0) {
              rt0 = TM - RT + ((warp & 1) * LPW + lane) * 4;
              if (lane < LPW) jt = *reinterpret_cast<const int4*>(ik + t * TM + rt0);
            }
            const uint32_t dst = a0 + (uint32_t)(slot * kABytes);
            if (t == 0 && tid == 0) TR(0, cnt[0] - 1);
            mbar_wait_warp(a_empty + slot, par ^ 1u);
            if (t == 0 && tid == 0) TR(1, cnt[0] - 1);
            if constexpr (RT > 0) {
              if (gt == 0) mbar_expect_tx(a_full + slot, RT * 128);
              if (lane < LPW)
                tma_gather4(dst + (uint32_t)(rt0 * 128), &map_feat, a_full + slot, c * KCH, jt.x, jt.y, jt.z, jt.w);
            }
#pragma unroll
            for (int i = 0; i < NR; ++i) {
              const bool v = (unsigned)j[i] < (unsigned)n_in;
              const int row = rb + i;
              cp_async16(dst + (uint32_t)(row * 128 + ((ch ^ (row & 7)) << 4)),
                         col + (v ? (size_t)(unsigned)j[i] * (unsigned)cin : 0), v ? 16u : 0u);
            }
            cp_async_arrive_noinc(a_full + slot);
            if (t == 0 && tid == 0) TR(2, cnt[0] - 1);
          }
        }
      }
    }
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp & 3;     // TMEM lane quadrant this warp may read
    const int part = warp >> 2; // the quadrant's (tile, 16-column) items are split over its 2 warps
    PH(3);
    if (any) {
      mbar_wait_warp(accum_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    PH(4);
    constexpr int kBlocks = CO / 16;
    for (int item = part; item < T * kBlocks; item += kProducerWarps / 4) {
      const int t = item / kBlocks, cb = (item % kBlocks) * 16;
      const long long o = (long long)(tile0 + t) * TM + q * 32 + lane;
      const bool live = ((t == 0 ? tmask[0] : t == 1 ? tmask[1] : t == 2 ? tmask[2] : tmask[3])) != 0;
      uint32_t v[16];
      if (live) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * C::kCols + cb);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
              "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
              "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
      }
      if (o < n_out && cb < cout) {
        float* dst = out + o * cout + cb;
        if (cb + 16 <= cout) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float4 r4 = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                                    __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
            if (bias) {
              const float4 b4 = ldg4(bias + cb + 4 * g);
              r4.x += b4.x; r4.y += b4.y; r4.z += b4.z; r4.w += b4.w;
            }
            *reinterpret_cast<float4*>(dst + 4 * g) = r4;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (cb + i < cout) dst[i] = __uint_as_float(v[i]) + (bias ? __ldg(bias + cb + i) : 0.f);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    PH(5);
  } else if (warp < kProducerWarps + kMaxT) {
    // ===================== MMA issuers: warp 8 + t drives tile t =====================
    const int t = warp - kProducerWarps;
    if (t < T) {
      // all 32 lanes walk the loop; values that come from shared memory are passed through a warp
      // reduction, whose result the compiler knows to be uniform
      constexpr uint32_t idesc = PREC ? make_idesc_bf16(CO) : make_idesc_tf32(CO);
      const uint32_t mt = __reduce_or_sync(0xffffffffu, t == 0 ? tmask[0] : t == 1 ? tmask[1] : t == 2 ? tmask[2] : tmask[3]);
      const uint32_t any_u = __reduce_or_sync(0xffffffffu, any);
      const uint32_t d_tmem = __reduce_or_sync(0xffffffffu, tmem_base) + (uint32_t)(t * C::kCols);
      const uint32_t a_ring = smem_u32(a_base) + (uint32_t)(t * ring * kABytes);
      int cnt = 0, sb = 0;    // stages of this tile consumed so far; filter-slice ring position
      uint32_t pb = 0, accumulate = 0;
      for (int k = 0; k < kvol; ++k) {
        if (!((any_u >> k) & 1u)) continue;
        const bool mine = (mt >> k) & 1u;
        for (int c = 0; c < n_chunks; ++c) {
          if (t == 0 && lane == 0 && mine) TR(3, cnt);
          mbar_wait_warp(b_full + sb, pb);
          if (mine) {
            const int slot = cnt & (ring - 1);
            if (t == 0 && lane == 0) TR(6, cnt);
            mbar_wait_warp(a_full + t * ring + slot, (uint32_t)(cnt >> ring_shift) & 1u);
            if (t == 0 && lane == 0) TR(4, cnt);
            ++cnt;
            if constexpr (!GATHER4) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (t == 0 && lane == 0) TR(7, cnt - 1);
            const uint64_t a_desc = make_desc_sw128(a_ring + (uint32_t)(slot * kABytes));
            const uint64_t b_desc = make_desc_sw128(smem_u32(b_base) + (uint32_t)(sb * C::kBBytes));
            if constexpr (PREC == 0) {
#pragma unroll
              for (int ks = 0; ks < KCH / 8; ++ks) {
                // 8 tf32 = 32 bytes along K inside the 128-byte swizzle row: +2 in 16-byte units
                umma_tf32_elect(d_tmem, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc, accumulate);
                accumulate = 1;
              }
            } else {
              // 16 bf16 = 32 bytes per K step: steps 0,1 = hi halves of the 32 channels, steps 2,3 = lo halves
#pragma unroll
              for (int term = 0; term < 3; ++term) {
                const int ao = term == 2 ? 4 : 0, bo = term == 1 ? 4 : 0;    // lo.hi last, hi.lo second
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  umma_bf16_elect(d_tmem, a_desc + (uint64_t)(ao + 2 * ks), b_desc + (uint64_t)(bo + 2 * ks), idesc,
                                  accumulate);
                  accumulate = 1;
                }
              }
            }
            if (t == 0 && lane == 0) TR(8, cnt - 1);
            umma_commit_elect(a_empty + t * ring + slot);
            umma_commit_elect(b_empty + sb);
            if (t == 0 && lane == 0) TR(5, cnt - 1);
          } else {
            // tile without this offset: release the filter slice (after it landed, so the arrival
            // is counted in the right phase)
            mbar_arrive_elect(b_empty + sb);
          }
          if (++sb == n_sb) { sb = 0; pb ^= 1u; }
        }
      }
      umma_commit_elect(accum_bar);
    }
    __syncwarp();
  } else {
    // ===================== B producer (tiled TMA) =====================
    if (lane == 0) {
      int sb = 0;
      uint32_t pb = 0;
      for (int k = 0; k < kvol; ++k) {
        if (!((any >> k) & 1u)) continue;
        for (int c = 0; c < n_chunks; ++c) {
          TR(9, tr_b);
          mbar_wait(b_empty + sb, pb ^ 1u);
          TR(10, tr_b);
#ifdef DDF_TRACE
          ++tr_b;
#endif
          mbar_expect_tx(b_full + sb, (uint32_t)(cout * KCH * 4));
          tma_tile_2d(smem_u32(b_base + sb * C::kBBytes), &map_w, b_full + sb, c * KCH, k * cout);
          if (++sb == n_sb) { sb = 0; pb ^= 1u; }
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
#ifdef DDF_PHASES
  if (tid == 0) {
    ph[6] = clock64();
    const unsigned slot = atomicAdd(&g_ph_n, 1u);
    if (slot < 16384) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      g_ph[slot][0] = smid;
      g_ph[slot][1] = blockIdx.x;
      for (int i = 0; i < 7; ++i) g_ph[slot][2 + i] = ph[i];
      g_ph[slot][9] = (unsigned long long)T;
    }
  }
#endif
#ifdef DDF_TRACE
  if (blockIdx.x == 3 && tid == 0) {
    const long long t0 = s_tr[0][0];
    for (int i = 0; i < 30; ++i)
      printf("stage %2d: prod wait %6lld slot %6lld issued %6lld | issuer start %6lld b_full %6lld a_full %6lld fenced %6lld mma %6lld committed %6lld | B wait %6lld empty %6lld\n", i,
             s_tr[0][i] - t0, s_tr[1][i] - t0, s_tr[2][i] - t0, s_tr[3][i] - t0, s_tr[6][i] - t0, s_tr[4][i] - t0,
             s_tr[7][i] - t0, s_tr[8][i] - t0, s_tr[5][i] - t0, s_tr[9][i] - t0, s_tr[10][i] - t0);
  }
#endif
  if (warp == kProducerWarps) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad, table driven:   gW[k] (Cin x Cout) = sum over output rows o of  feat[G[o, k], :]^T . gout[o, :]
//
// The pair-list kernel (sparse_conv_tc.cu) gathers BOTH operands and walks one offset at a time, i.e.
// 27 streaming passes over feat and gout (2.3 GB of DRAM reads for the 64-channel stage of the bench
// workload, ncu).  Here the output rows are walked ONCE: a CTA owns a group of KG = 512 / Cout kernel
// offsets (one TMEM accumulator each) and a range of 32-row sub-tiles; per sub-tile the gout rows
// (dense, contiguous) and the 32 x K slice of the gather table are staged once and reused by the KG
// offsets, only the feat rows are gathered (neighbours of consecutive output rows: L2 hits).
// GEMM per stage: M = Cin (128 TMEM lanes, lanes >= Cin unused), N = Cout, K = 32 rows; both
// operands are MN-major, SWIZZLE_128B_BASE32B: atoms of 4 rows x 32 channels (512 bytes), the
// 32-byte chunk index XORed with (row & 3).
// Warp roles as in the forward kernel: warps 0..7 gather A in 4 groups (group = slot & 3) and drain
// the accumulators at the end (red.global.add.v4), warps 8..11 issue the MMAs of the offsets
// kk = j (mod 4), warp 12 stages gout + table.
// ------------------------------------------------------------------------------------------------
constexpr int kWgSlotsA = 8;        // 2 per issuer
constexpr int kWgSlotsB = 3;
// rows per sub-tile (= K extent of a stage) is a template parameter; 32 is what ships

__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(512 >> 4) << 16;           // LBO: next 32-channel block
  d |= (uint64_t)(sbo_bytes >> 4) << 32;     // SBO: next 4-row group
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                    // SWIZZLE_128B_BASE32B
  return d;
}
// byte offset of 16-byte chunk `ch` (0..7) of row `p` in channel block `mb` (nblk blocks per 4-row group)
__device__ __forceinline__ uint32_t mn_chunk_offset(int p, int mb, int ch, int nblk) {
  const int r = p & 3;
  return (uint32_t)((((p >> 2) * nblk + mb) << 9) + (r << 7) + ((((ch >> 1) ^ r)) << 5) + ((ch & 1) << 4));
}

// 16 producer warps = 8 groups, one per A slot: what bounds the kernel is how many producer groups issue gathers
// concurrently on an SM (in-kernel trace: a group needs about 2000 cycles per 16 KB stage, the MMA issuers wait for
// data), and the rings of this kernel only leave room for one CTA per SM.
constexpr int kWgProducerWarps = 16;
constexpr int kWgThreads = (kWgProducerWarps + kMaxT + 2) * 32;   // + 4 MMA issuer warps + 2 staging warps

// PK = kernel offsets packed into the M dimension of one MMA (PK * Cin = 128 TMEM lanes): with 32 or 64 channels an
// unpacked accumulator uses a quarter / half of the lanes and the issuers spend their time issuing 4x / 2x the MMAs
// (about 100 cycles each: in-kernel trace, profiles/r2_conv_analysis.md).  A packed A stage is laid out like the A
// stage of a 128-channel layer whose 32-channel blocks come from PK different gathered rows.
template <int CO, int SR, int PK>
__global__ void __launch_bounds__(kWgThreads)
spconv_wgrad_table_kernel(const float* __restrict__ feat, const float* __restrict__ gout,
                          const int* __restrict__ table, float* __restrict__ gw, int n_out, int n_in,
                          int kvol, int cin, int cout, int KG, int G) {
  ddf::pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int kTblBytes = (SR * kMaxKvol * 4 + 1023) / 1024 * 1024;
  const int a_nb = (PK * cin) >> 5, b_nb = CO >> 5;    // 32-channel blocks per row (cin == CO when PK > 1)
  const int a_bytes = SR * PK * cin * 4, b_bytes = SR * CO * 4;
  uint8_t* a_base = smem;
  uint8_t* b_base = a_base + kWgSlotsA * a_bytes;
  uint8_t* t_base = b_base + kWgSlotsB * b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(t_base + kWgSlotsB * kTblBytes + 2048);  // + over-read slack
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kWgSlotsA;
  uint64_t* b_full = a_empty + kWgSlotsA;
  uint64_t* b_empty = b_full + kWgSlotsB;
  uint64_t* accum_bar = b_empty + kWgSlotsB;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef DDF_TRACE
  __shared__ long long w_tr[8][64];
  int tr_p = 0, tr_i = 0, tr_b = 0;
#define WTR(ev, i) do { if (blockIdx.x == 5 && (i) < 64) w_tr[ev][i] = clock64(); } while (0)
#else
#define WTR(ev, i) do {} while (0)
#endif
  const int g = blockIdx.x % G, split = blockIdx.x / G, S = gridDim.x / G;
  const int k0 = g * KG;
  const int nko = min(KG, kvol - k0);          // kernel offsets of this CTA
  const int nk = (nko + PK - 1) / PK;          // packed units (accumulators) of this CTA
  const int NS = (n_out + SR - 1) / SR;
  const int st_begin = (int)((long long)NS * split / S), st_end = (int)((long long)NS * (split + 1) / S);
  constexpr uint32_t kCols = CO;     // CO >= 32 here
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)nk * kCols) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < kWgSlotsA; ++s) {
      mbar_init(a_full + s, 64);
      mbar_init(a_empty + s, 1);
    }
    for (int s = 0; s < kWgSlotsB; ++s) {
      mbar_init(b_full + s, 64);
      mbar_init(b_empty + s, kMaxT);
    }
    mbar_init(accum_bar, kMaxT);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWgProducerWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  const bool work = st_begin < st_end;

  if (warp < kWgProducerWarps) {
    // ===================== A producers =====================
    if (work) {
      const int grp = warp >> 1, gt = tid & 63;
      constexpr int kChunksO = CO / 4;                   // 16-byte chunks per gathered row (Cin == Cout == CO)
      constexpr int kChunks = PK * kChunksO;             // ... per row of the packed stage
      constexpr int kRowsPerPass = 64 / kChunks;         // rows covered by the 64 threads of a group
      constexpr int kIters = SR / kRowsPerPass;          // chunks per thread and stage
      const int c32 = gt % kChunks, pr = gt / kChunks;   // this thread's chunk of the packed row and first row
      const int jo = c32 / kChunksO, cc = c32 % kChunksO;   // offset inside the unit, chunk of the gathered row
      int cnt[kMaxT] = {0, 0, 0, 0};
      int sb = 0;
      uint32_t pb = 0;
      for (int st = st_begin; st < st_end; ++st) {
        bool have_tbl = false;
        const int* tbl = reinterpret_cast<const int*>(t_base + sb * kTblBytes);
        const int rows_left = n_out - st * SR;
#pragma unroll 1
        for (int kk = 0; kk < nk; ++kk) {
          const int j = kk & (kMaxT - 1);
          int c;
          if (j == 0) c = cnt[0]++; else if (j == 1) c = cnt[1]++; else if (j == 2) c = cnt[2]++; else c = cnt[3]++;
          const int sl = j * 2 + (c & 1);
          if (sl != grp) continue;
          if (!have_tbl) {
            mbar_wait_warp(b_full + sb, pb);     // table slice (and gout rows) of this sub-tile landed
            have_tbl = true;
          }
          const int k = k0 + kk * PK + jo;
          const bool kv = kk * PK + jo < nko;            // the last unit of a CTA may be partly empty: zero rows
          int rowidx[kIters];
#pragma unroll
          for (int i = 0; i < kIters; ++i) rowidx[i] = kv ? tbl[(pr + i * kRowsPerPass) * kvol + k] : -1;
#ifdef DDF_TRACE
          if (tid == 0) WTR(0, tr_p);
#endif
          mbar_wait_warp(a_empty + sl, (uint32_t)((c >> 1) & 1) ^ 1u);
#ifdef DDF_TRACE
          if (tid == 0) WTR(1, tr_p);
#endif
          const uint32_t dst = smem_u32(a_base + sl * a_bytes);
          const float* src0 = feat + cc * 4;
#pragma unroll
          for (int i = 0; i < kIters; ++i) {
            const int p = pr + i * kRowsPerPass;
            const bool v = rowidx[i] >= 0 && p < rows_left;
            cp_async16(dst + mn_chunk_offset(p, c32 >> 3, c32 & 7, a_nb),
                       src0 + (v ? (size_t)(unsigned)rowidx[i] * (unsigned)CO : 0), v ? 16u : 0u);
          }
          cp_async_arrive_noinc(a_full + sl);
#ifdef DDF_TRACE
          if (tid == 0) { WTR(2, tr_p); ++tr_p; }
#endif
        }
        if (++sb == kWgSlotsB) { sb = 0; pb ^= 1u; }
      }
    }
    // ===================== epilogue: accumulators -> gW (red.global.add) =====================
    if (work) {
      mbar_wait_warp(accum_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int q = warp & 3, half = warp >> 2;
      const int lrow = q * 32 + lane;                    // TMEM lane = (offset inside the unit, input channel)
      const int jl = PK > 1 ? lrow / CO : 0, ci = PK > 1 ? lrow % CO : lrow;
      for (int kk = half; kk < nk; kk += kWgProducerWarps / 4) {
        const bool lane_live = ci < cin && kk * PK + jl < nko;
        for (int cb = 0; cb < cout; cb += 16) {
          uint32_t v[16];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kk * kCols + cb);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
                "=r"(v[14]), "=r"(v[15])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (lane_live) {
            float* dst = gw + ((long long)(k0 + kk * PK + jl) * cin + ci) * cout + cb;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              red_add_v4(dst + 4 * i, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                         __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
  } else if (warp < kWgProducerWarps + kMaxT) {
    // ===================== MMA issuers: offsets kk = j (mod 4) =====================
    const int j = warp - kWgProducerWarps;
    if (work) {
      constexpr uint32_t idesc = make_idesc_tf32(CO) | (1u << 15) | (1u << 16);   // A, B MN-major
      const uint32_t tm = __reduce_or_sync(0xffffffffu, tmem_base);
      const uint32_t a_ring = smem_u32(a_base) + (uint32_t)(j * 2 * a_bytes);
      int cnt = 0, sb = 0;
      uint32_t pb = 0;
      for (int st = st_begin; st < st_end; ++st) {
        mbar_wait_warp(b_full + sb, pb);
        const uint32_t b_smem = smem_u32(b_base) + (uint32_t)(sb * b_bytes);
        bool issued = false;
        for (int kk = j; kk < nk; kk += kMaxT) {
          const int sl = cnt & 1;
#ifdef DDF_TRACE
          if (j == 0 && lane == 0) WTR(3, tr_i);
#endif
          mbar_wait_warp(a_full + j * 2 + sl, (uint32_t)(cnt >> 1) & 1u);
#ifdef DDF_TRACE
          if (j == 0 && lane == 0) WTR(4, tr_i);
#endif
          ++cnt;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_smem = a_ring + (uint32_t)(sl * a_bytes);
#pragma unroll
          for (int ks = 0; ks < SR / 8; ++ks) {
            // one MMA (K = 8 rows) spans two 4-row groups
            const uint64_t a_desc = make_desc_mn_sw128(a_smem + ks * 2 * a_nb * 512, a_nb * 512);
            const uint64_t b_desc = make_desc_mn_sw128(b_smem + ks * 2 * b_nb * 512, b_nb * 512);
            umma_tf32_elect(tm + (uint32_t)kk * kCols, a_desc, b_desc, idesc, (st != st_begin || ks != 0) ? 1u : 0u);
          }
          umma_commit_elect(a_empty + j * 2 + sl);
#ifdef DDF_TRACE
          if (j == 0 && lane == 0) { WTR(5, tr_i); ++tr_i; }
#endif
          issued = true;
        }
        if (issued) umma_commit_elect(b_empty + sb); else mbar_arrive_elect(b_empty + sb);
        if (++sb == kWgSlotsB) { sb = 0; pb ^= 1u; }
      }
      umma_commit_elect(accum_bar);
    }
    __syncwarp();
  } else {
    // ===================== gout rows + table slice of each sub-tile (cp.async, 2 warps) =====================
    if (work) {
      constexpr int kChunks = CO / 4;
      const int bt = tid - (kWgProducerWarps + kMaxT) * 32;     // 0..63
      int sb = 0;
      uint32_t pb = 0;
      const long long tbl_bytes_total = (long long)n_out * kvol * 4;
      const int tbl_chunks = (SR * kvol * 4 + 15) / 16;
      for (int st = st_begin; st < st_end; ++st) {
#ifdef DDF_TRACE
        if (bt == 0) WTR(6, tr_b);
#endif
        mbar_wait_warp(b_empty + sb, pb ^ 1u);
        const uint32_t tdst = smem_u32(t_base + sb * kTblBytes);
        const long long tsrc = (long long)st * SR * kvol * 4;
#pragma unroll 4
        for (int e = bt; e < tbl_chunks; e += 64) {
          const long long off = tsrc + e * 16;
          const long long left = tbl_bytes_total - off;
          const uint32_t nb = left >= 16 ? 16u : (left > 0 ? (uint32_t)left : 0u);
          cp_async16(tdst + e * 16, reinterpret_cast<const uint8_t*>(table) + (nb ? off : 0), nb);
        }
        const uint32_t bdst = smem_u32(b_base + sb * b_bytes);
        const long long row0 = (long long)st * SR;
#pragma unroll
        for (int i = 0; i < SR * kChunks / 64; ++i) {
          const int e = bt + i * 64;
          const int p = e / kChunks, cc = e % kChunks;
          const bool v = row0 + p < n_out;
          cp_async16(bdst + mn_chunk_offset(p, cc >> 3, cc & 7, b_nb), gout + (v ? (row0 + p) * CO + cc * 4 : 0), v ? 16u : 0u);
        }
        cp_async_arrive_noinc(b_full + sb);
#ifdef DDF_TRACE
        if (bt == 0) { WTR(7, tr_b); ++tr_b; }
#endif
        if (++sb == kWgSlotsB) { sb = 0; pb ^= 1u; }
      }
    }
  }
  __syncthreads();
#ifdef DDF_TRACE
  if (blockIdx.x == 5 && tid == 0) {
    const long long t0 = w_tr[6][0];
    for (int i = 0; i < 24; ++i)
      printf("wgrad %2d: prod wait %6lld slot %6lld issued %6lld | issuer wait %6lld full %6lld commit %6lld | B wait %6lld loaded %6lld\n", i,
             w_tr[0][i] - t0, w_tr[1][i] - t0, w_tr[2][i] - t0, w_tr[3][i] - t0, w_tr[4][i] - t0, w_tr[5][i] - t0, w_tr[6][i] - t0, w_tr[7][i] - t0);
  }
#endif
  if (warp == kWgProducerWarps) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols)
                 : "memory");
  }
}

template <int CO>
int launch_wgrad_table(const float* feat, const float* gout, const int* table, float* gw, int64_t n_out,
                       int64_t n_in, int kvol, int cin, int cout, cudaStream_t stream) {
  // 32 output rows per stage; 128 / CO kernel offsets share one MMA (M = 128 lanes), so an A stage is always
  // 32 rows x 512 bytes = 16 KB
  constexpr int SR = 32;
  constexpr int PK = 128 / CO;
  constexpr int kTblBytes = (SR * kMaxKvol * 4 + 1023) / 1024 * 1024;
  const int smem = kWgSlotsA * SR * PK * cin * 4 + kWgSlotsB * SR * CO * 4 + kWgSlotsB * kTblBytes + 2048 + 512 + 1024;
  DDF_SET_SMEM_ONCE((spconv_wgrad_table_kernel<CO, SR, PK>), smem);
  // two CTAs per SM when the rings are small; they then share the 512 TMEM columns
  const int per_sm = smem <= 110 * 1024 ? 2 : 1;
  int KG = (512 / per_sm) / CO * PK;    // kernel offsets per CTA = accumulators x offsets per accumulator
  if (KG > kvol) KG = kvol;
  const int G = (kvol + KG - 1) / KG;
  const long long NS = ddf::cdiv(n_out, SR);
  long long S = (ddf::kNumSM * per_sm) / G;
  if (S > NS) S = NS;
  if (S < 1) S = 1;
  DDF_LAUNCH_PDL((spconv_wgrad_table_kernel<CO, SR, PK>), (unsigned)(G * S), kWgThreads, smem, stream, feat, gout, table, gw,
             (int)n_out, (int)n_in, kvol, cin, cout, KG, G);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// ---- host side: tensor maps ---------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 2-D fp32 row-major [rows, cols] tensor, box [box_rows x 32 floats], 128-byte swizzle, zero OOB fill
bool make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)(rows > 0 ? rows : 1)};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {(cuuint32_t)KCH, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int CO, bool GATHER4, int PREC, int RT>
int launch_tma(const float* feat, const float* wt, const int* table, const float* bias, float* out,
               int64_t n_out, int64_t n_in, int kvol, int cin, int cout, cudaStream_t stream) {
  using C = Cfg<CO>;
  DDF_SET_SMEM_ONCE((spconv_tma_kernel<CO, GATHER4, PREC, RT>), C::kSmemBytes);
  CUtensorMap map_feat, map_w;
  if (!make_map(&map_feat, feat, n_in, cin, 1) || !make_map(&map_w, wt, (int64_t)kvol * cout, cin, cout)) {
    ddf::set_error("sparse conv: cuTensorMapEncodeTiled failed (n_in=%lld cin=%d cout=%d)", (long long)n_in, cin, cout);
    return DDF_ERR_CUDA;
  }
  const long long ntiles = ddf::cdiv(n_out, TM);
  // tiles per CTA: enough CTAs to fill the SMs first, then share filter slices between tiles
  int T = (int)ddf::cdiv(ntiles, ddf::kNumSM);
  const int tmax = 512 / C::kCols < kMaxT ? 512 / C::kCols : kMaxT;
  if (T > tmax) T = tmax;
  if (T < 1) T = 1;
  int n_slots = kStagesA, n_sb = C::kStagesB;
  if (CO <= 64 && ntiles >= 4 * ddf::kNumSM) {
    // narrow output, many tiles: two CTAs per SM (2 tiles, 4 A slots, 2 filter slices each) so that the
    // table load and the epilogue of one CTA overlap the main loop of the other
    T = 2;
    n_slots = 4;
    n_sb = 2;
  }
#if defined(DDF_TRACE) || defined(DDF_TUNE)
  // instrumented builds only: schedule overrides for A/B timing
  if (const char* e = getenv("DDF_TMA_T")) { T = atoi(e); if (T > tmax) T = tmax; }
  if (const char* e = getenv("DDF_TMA_SLOTS")) n_slots = atoi(e);
  if (const char* e = getenv("DDF_TMA_SB")) n_sb = atoi(e);
#endif
  const int smem = n_slots * kABytes + n_sb * C::kBBytes + T * TM * kMaxKvol * 4 + 512 + 1024;
  const unsigned grid = (unsigned)ddf::cdiv(ntiles, T);
  DDF_LAUNCH_PDL((spconv_tma_kernel<CO, GATHER4, PREC, RT>), grid, kThreads, smem, stream, map_feat, map_w, feat, table,
             bias, out, (int)n_out, (int)n_in, kvol, cin, cout, T, n_slots, n_sb);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

}  // namespace

namespace ddf {

// the TMA kernel takes full 32-channel chunks (narrower layers stay on the cp.async kernel) and
// output widths that are a multiple of 16
bool spconv_tma_supported(int kvol, int cin, int cout) {
  return encode_fn() != nullptr && kvol <= kMaxKvol && cin % KCH == 0 && cout % 16 == 0 && cout >= 16 && cout <= 128;
}

// feat [n_in, cin]; wt [K, cout, cin] (K-major B operand); table [n_out, K].  split = both operands are in the
// bf16 hi/lo block layout (see the kernel comment), else fp32.
int spconv_tma_launch(const float* feat, const float* wt, const int* table, const float* bias, float* out,
                      int64_t n_out, int64_t n_in, int kvol, int cin, int cout, bool gather4, bool split,
                      cudaStream_t stream) {
  // rows of every A stage that go through TMA gather4 instead of LDGSTS (hybrid gather; 0 = LDGSTS only)
  static const int rows_tma = [] {
    const char* e = getenv("DDF_CONV_ROWS_TMA");
    const int v = e ? atoi(e) : kDefaultRowsTma;
    return v >= 64 ? 64 : v >= 32 ? 32 : 0;
  }();
#define DDF_TMA_ARGS feat, wt, table, bias, out, n_out, n_in, kvol, cin, cout, stream
#define DDF_TMA_CASE(CO)                                                                   \
  if (gather4 && !split) return launch_tma<CO, true, 0, 0>(DDF_TMA_ARGS);                  \
  if (split) {                                                                             \
    if (rows_tma == 64) return launch_tma<CO, false, 1, 64>(DDF_TMA_ARGS);                 \
    if (rows_tma == 32) return launch_tma<CO, false, 1, 32>(DDF_TMA_ARGS);                 \
    return launch_tma<CO, false, 1, 0>(DDF_TMA_ARGS);                                      \
  }                                                                                        \
  if (rows_tma == 64) return launch_tma<CO, false, 0, 64>(DDF_TMA_ARGS);                   \
  if (rows_tma == 32) return launch_tma<CO, false, 0, 32>(DDF_TMA_ARGS);                   \
  return launch_tma<CO, false, 0, 0>(DDF_TMA_ARGS)
  if (cout <= 16) { DDF_TMA_CASE(16); }
  if (cout <= 32) { DDF_TMA_CASE(32); }
  if (cout <= 64) { DDF_TMA_CASE(64); }
  DDF_TMA_CASE(128);
#undef DDF_TMA_CASE
#undef DDF_TMA_ARGS
}

// table-driven wgrad: SubM layers with Cin == Cout in {32, 64, 128} (the sub-tile of the gather table
// must fit its smem slot: kvol <= 27)
bool spconv_wgrad_table_supported(int kvol, int cin, int cout) {
  return kvol <= kMaxKvol && cin == cout && (cin == 32 || cin == 64 || cin == 128);
}

// gw [K, cin, cout] must be zeroed by the caller; table [n_out, K]
int spconv_wgrad_table_launch(const float* feat, const float* gout, const int* table, float* gw,
                              int64_t n_out, int64_t n_in, int kvol, int cin, int cout, cudaStream_t stream) {
  if (cout <= 32) return launch_wgrad_table<32>(feat, gout, table, gw, n_out, n_in, kvol, cin, cout, stream);
  if (cout <= 64) return launch_wgrad_table<64>(feat, gout, table, gw, n_out, n_in, kvol, cin, cout, stream);
  return launch_wgrad_table<128>(feat, gout, table, gw, n_out, n_in, kvol, cin, cout, stream);
}

}  // namespace ddf
