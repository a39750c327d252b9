// Sparse 3-D convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05, tf32 inputs,
// fp32 accumulation in TMEM) for sm_100a.  Used for forward and dgrad when the contraction width
// (input channels of the launch) is a multiple of 32; narrower layers stay on the SIMT kernel.
//
//   out[o, :] = sum_k  feat[table[o, k], :] . Wt[k]^T         Wt[k] : [CO rows (out ch), CI (in ch)]
//
// CTA = 128 output rows (the MMA M dimension, one TMEM lane per row) x all output channels (MMA N,
// 16..128 TMEM columns).  The contraction runs over (kernel offset k, 32-channel chunk): one
// pipeline stage = A tile [128 rows x 32 floats] gathered through the rulebook table + B tile
// [CO x 32 floats] of the filter slice, both K-major with the 128-byte swizzle the UMMA smem
// descriptors expect, written by cp.async (zero-fill for missing neighbours) and tracked by
// mbarriers; four tcgen05.mma (K = 8 each) consume a stage.  Offsets for which no row of the tile
// has a neighbour are skipped (per-tile 32-bit mask).
//
// Warp roles: warps 0-3 gather (1 row-chunk stream each) and later run the epilogue
// (tcgen05.ld 32 lanes x 32 columns -> +bias -> global); warp 4 allocates TMEM and its elected lane
// issues the MMAs and commits.
#include "common.cuh"

namespace {

constexpr int TM = 128;           // rows per CTA = UMMA M
constexpr int KCH = 32;           // floats per K chunk (128 bytes = one swizzle row)
constexpr int kProducers = 128;
constexpr int kThreadsTC = 160;
constexpr int kMaxKvol = 27;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
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
// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ constexpr uint32_t make_idesc_tf32(int n) {
  return (1u << 4)                      // D format f32
         | (2u << 7) | (2u << 10)       // A, B format tf32
         | ((uint32_t)(n >> 3) << 17)   // N
         | ((uint32_t)(TM >> 4) << 24); // M
}

template <int CO>
struct TcCfg {
  static constexpr int kStages = CO >= 128 ? 3 : 4;
  static constexpr int kABytes = TM * KCH * 4;       // 16 KB
  static constexpr int kBBytes = CO * KCH * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = CO < 32 ? 32 : CO;
  // stages + table cache + barriers (+1 KB alignment slack)
  static constexpr int kSmemBytes = kStages * kStageBytes + TM * 27 * 4 + 256 + 1024;  // table cache sized for kvol <= 27
};

template <int CO>
__global__ void __launch_bounds__(kThreadsTC)
spconv_tc_kernel(const float* __restrict__ feat, const float* __restrict__ wt,
                 const int* __restrict__ table, const float* __restrict__ bias,
                 float* __restrict__ out, int n_out, int kvol, int cin, int cout) {
  ddf::pdl_sync();
  using Cfg = TcCfg<CO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  int* s_idx = reinterpret_cast<int*>(smem + Cfg::kStages * Cfg::kStageBytes);  // [TM][kvol]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_idx + TM * kMaxKvol);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
  uint64_t* accum_bar = bars + 2 * Cfg::kStages;   // [1]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 1);
  uint32_t* s_mask = s_tmem + 1;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * TM;

  if (tid == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar + s, kProducers);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(accum_bar, 1);
    *s_mask = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(s_tmem)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  // table rows of this tile -> smem, active-offset mask
  if (tid < kProducers) {
    const int o = row0 + tid;
    uint32_t m = 0;
    for (int k = 0; k < kvol; ++k) {
      const int j = o < n_out ? __ldg(table + (long long)o * kvol + k) : -1;
      s_idx[tid * kvol + k] = j;
      if (j >= 0) m |= 1u << k;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, off);
    if (lane == 0 && m) atomicOr(s_mask, m);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t mask = *s_mask;
  const uint32_t tmem_base = *s_tmem;
  const int n_chunks = (cin + KCH - 1) / KCH;
  const int kc = cin < KCH ? cin : KCH;   // floats per chunk actually present (cin < 32: one partial chunk)
  const int live_chunks = kc >> 2;         // 16-byte pieces per row to copy

  if (warp < 4) {
    // ===================== producers: gather A, stream B =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int k = 0; k < kvol; ++k) {
      if (!((mask >> k) & 1u)) continue;
      const float* wk = wt + (long long)k * cout * cin;
      for (int c = 0; c < n_chunks; ++c) {
        mbar_wait(empty_bar + stage, phase ^ 1u);
        const uint32_t a_smem = smem_u32(stage_base + stage * Cfg::kStageBytes);
        const uint32_t b_smem = a_smem + Cfg::kABytes;
        const int c0 = c * KCH;
#pragma unroll
        for (int i = 0; i < (TM * 8) / kProducers; ++i) {
          const int e = i * kProducers + tid;
          const int r = e >> 3, ch = e & 7;
          if (ch < live_chunks) {
            const int j = s_idx[r * kvol + k];
            const float* src = feat + (j >= 0 ? ((long long)j * cin + c0 + ch * 4) : 0);
            cp_async16(a_smem + r * 128 + ((ch ^ (r & 7)) << 4), src, j >= 0 ? 16u : 0u);
          }
        }
#pragma unroll
        for (int i = 0; i < (CO * 8 + kProducers - 1) / kProducers; ++i) {
          const int e = i * kProducers + tid;
          if (e < CO * 8) {
            const int r = e >> 3, ch = e & 7;
            if (ch < live_chunks) {
              const bool ok = r < cout;
              const float* src = wk + (ok ? ((long long)r * cin + c0 + ch * 4) : 0);
              cp_async16(b_smem + r * 128 + ((ch ^ (r & 7)) << 4), src, ok ? 16u : 0u);
            }
          }
        }
        cp_async_arrive_noinc(full_bar + stage);
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int o = row0 + warp * 32 + lane;
    if (mask) {
      mbar_wait(accum_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
#pragma unroll
    for (int cb = 0; cb < CO; cb += 16) {
      uint32_t v[16];
      if (mask) {
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb;
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
      if (o < n_out) {
        float* dst = out + (long long)o * cout + cb;
        if (cb + 16 <= cout && (cout & 3) == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 r4 = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                    __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
            if (bias) {
              r4.x += __ldg(bias + cb + 4 * q);
              r4.y += __ldg(bias + cb + 4 * q + 1);
              r4.z += __ldg(bias + cb + 4 * q + 2);
              r4.w += __ldg(bias + cb + 4 * q + 3);
            }
            *reinterpret_cast<float4*>(dst + 4 * q) = r4;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (cb + i < cout) dst[i] = __uint_as_float(v[i]) + (bias ? __ldg(bias + cb + i) : 0.f);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ===================== MMA issuer =====================
    if (lane == 0 && mask) {
      constexpr uint32_t idesc = make_idesc_tf32(CO);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accumulate = 0;
      for (int k = 0; k < kvol; ++k) {
        if (!((mask >> k) & 1u)) continue;
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait(full_bar + stage, phase);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_smem = smem_u32(stage_base + stage * Cfg::kStageBytes);
          const uint64_t a_desc = make_desc_sw128(a_smem);
          const uint64_t b_desc = make_desc_sw128(a_smem + Cfg::kABytes);
          for (int ks = 0; ks < (kc >> 3); ++ks) {
            // only the K steps backed by real channels are issued; advance 8 tf32 = 32 bytes along K inside the 128-byte swizzle row: +2 in 16-byte units
            umma_tf32(tmem_base, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(empty_bar + stage);  // frees the smem stage when these MMAs retire
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
  }
}

template <int CO>
int launch_tc(const float* feat, const float* wt, const int* table, const float* bias, float* out,
              int64_t n_out, int kvol, int cin, int cout, cudaStream_t stream) {
  using Cfg = TcCfg<CO>;
  DDF_SET_SMEM_ONCE(spconv_tc_kernel<CO>, Cfg::kSmemBytes);
  const unsigned grid = (unsigned)ddf::cdiv(n_out, TM);
  DDF_LAUNCH_PDL(spconv_tc_kernel<CO>, grid, kThreadsTC, Cfg::kSmemBytes, stream, feat, wt, table, bias,
             out, (int)n_out, kvol, cin, cout);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}


// ------------------------------------------------------------------------------------------------
// wgrad on tensor cores:  gW[k] (Cin x Cout) = sum over the pairs (i, o) of offset k of
//   feat[i, :]^T . gout[o, :]
// = a GEMM with M = Cin (128 TMEM lanes; lanes >= Cin hold don't-care rows), N = Cout, K = pairs.
// The gathered rows are channel-contiguous, i.e. both operands are MN-major.  For 32-bit MN-major
// operands the only UMMA smem layout is SWIZZLE_128B_BASE32B: atoms of 4 pairs x 32 channels
// (4 rows of 128 bytes), the 32-byte chunk index XORed with (pair & 3).  Atoms of one 4-pair group
// are contiguous (LBO = 512 B between 32-channel blocks, SBO between 4-pair groups); one tf32 MMA
// (K = 8) spans two groups.
//
// Scheduling: the pair lists of all kernel offsets are cut into stages of WP pairs and the global
// stage sequence (offset-major) is divided EVENLY over a persistent grid of 2 CTAs per SM, so the
// load is balanced whatever the per-offset pair counts are.  A CTA accumulates a segment (its
// stages of one offset) in TMEM, adds the tile to gW[k] with 16-byte red.global, and moves on to
// the next offset if its range crosses one.  Pair indices of stage g+1 are fetched while stage g
// is being issued (one index per producer thread, staged through shared memory).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(512 >> 4) << 16;           // LBO: next 32-channel block
  d |= (uint64_t)(sbo_bytes >> 4) << 32;     // SBO: next 4-pair group
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                    // SWIZZLE_128B_BASE32B
  return d;
}
// byte offset of 16-byte chunk `ch` (0..7) of pair `p` in channel block `mb` (nblk blocks per group)
__device__ __forceinline__ uint32_t mn_chunk_offset(int p, int mb, int ch, int nblk) {
  const int r = p & 3;
  return (uint32_t)((((p >> 2) * nblk + mb) << 9) + (r << 7) + ((((ch >> 1) ^ r)) << 5) + ((ch & 1) << 4));
}

constexpr int kWgStages = 3;
constexpr int kWgStageBytes = 32 * 1024;                       // A + B tile budget per stage
constexpr int kWgSmemBytes = kWgStages * kWgStageBytes + 2048 /*dead-block over-read slack*/ +
                             2 * 128 * 4 /*pair indices*/ + 512 /*barriers, offsets*/ + 1024;

template <int CO>
__global__ void __launch_bounds__(kThreadsTC)
spconv_wgrad_tc_kernel(const float* __restrict__ feat, const float* __restrict__ gout,
                       const int* __restrict__ pairs, const int* __restrict__ num, int pair_stride,
                       int kvol, int cin, int cout, int inverse, float* __restrict__ gw) {
  ddf::pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  int* s_pidx = reinterpret_cast<int*>(smem + kWgStages * kWgStageBytes + 2048);  // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_pidx + 256);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kWgStages;
  uint64_t* accum_bar = bars + 2 * kWgStages;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kWgStages + 1);
  int* s_first = reinterpret_cast<int*>(s_tmem + 1);  // [kvol + 1] first global stage of each offset

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kTmemCols = CO < 32 ? 32 : CO;
  const int a_nb = (cin + 31) >> 5;            // 32-channel blocks per pair row
  const int b_nb = CO < 32 ? 1 : CO / 32;
  // pairs per stage: 64 when two tiles of 64 rows fit the stage budget, else 32
  const int WP = (64 * (a_nb + b_nb) * 128 <= kWgStageBytes) ? 64 : 32;
  const int a_bytes = WP * a_nb * 128;

  if (tid == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(full_bar + s, kProducers);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(accum_bar, 1);
    int acc = 0;
    for (int k = 0; k < kvol; ++k) {
      s_first[k] = acc;
      acc += (num[k] + WP - 1) / WP;
    }
    s_first[kvol] = acc;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  const long long total = s_first[kvol];
  const int g_begin = (int)(total * blockIdx.x / gridDim.x);
  const int g_end = (int)(total * (blockIdx.x + 1) / gridDim.x);

  if (g_begin < g_end) {
    int k = 0;
    while (s_first[k + 1] <= g_begin) ++k;  // offset that owns the first stage
    if (warp < 4) {
      // ===================== producers (+ epilogue at segment ends) =====================
      int stage = 0;
      uint32_t phase = 0, accum_phase = 0;
      const int a_chunks = cin >> 2, b_chunks = cout >> 2;
      // which pair index this thread fetches for a stage: threads 0..63 input rows, 64..127 output rows
      auto fetch = [&](int g, int kk) -> int {
        const int p = (g - s_first[kk]) * WP + (tid & 63);
        if ((tid & 63) >= WP || p >= num[kk]) return -1;
        const int col = (tid < 64) ? (inverse ? 1 : 0) : (inverse ? 0 : 1);
        return __ldg(pairs + ((long long)kk * 2 + col) * pair_stride + p);
      };
      s_pidx[tid] = fetch(g_begin, k);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      int g = g_begin;
      while (g < g_end) {
        const int seg_end = min(g_end, s_first[k + 1]);
        for (; g < seg_end; ++g) {
          // prefetch the next stage's pair index (possibly of the next offset)
          int nk = k;
          if (g + 1 >= s_first[k + 1]) {
            nk = k + 1;
            while (nk < kvol && s_first[nk + 1] <= g + 1) ++nk;
          }
          const int nxt = (g + 1 < g_end) ? fetch(g + 1, nk) : -1;
          const int* idx_a = s_pidx + ((g - g_begin) & 1) * 128;
          const int* idx_b = idx_a + 64;
          mbar_wait(empty_bar + stage, phase ^ 1u);
          const uint32_t a_smem = smem_u32(smem + stage * kWgStageBytes);
          const uint32_t b_smem = a_smem + a_bytes;
          for (int e = tid; e < WP * a_chunks; e += kProducers) {
            const int p = e / a_chunks, cc = e % a_chunks;
            const int row = idx_a[p];
            cp_async16(a_smem + mn_chunk_offset(p, cc >> 3, cc & 7, a_nb),
                       feat + (row >= 0 ? ((long long)row * cin + cc * 4) : 0), row >= 0 ? 16u : 0u);
          }
          for (int e = tid; e < WP * b_chunks; e += kProducers) {
            const int p = e / b_chunks, cc = e % b_chunks;
            const int row = idx_b[p];
            cp_async16(b_smem + mn_chunk_offset(p, cc >> 3, cc & 7, b_nb),
                       gout + (row >= 0 ? ((long long)row * cout + cc * 4) : 0), row >= 0 ? 16u : 0u);
          }
          cp_async_arrive_noinc(full_bar + stage);
          if (++stage == kWgStages) {
            stage = 0;
            phase ^= 1u;
          }
          s_pidx[((g + 1 - g_begin) & 1) * 128 + tid] = nxt;
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        // segment done: drain the accumulator of offset k into gW[k]
        mbar_wait(accum_bar, accum_phase);
        accum_phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int ci = warp * 32 + lane;
#pragma unroll
        for (int cb = 0; cb < CO; cb += 16) {
          uint32_t v[16];
          const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb;
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
                "=r"(v[14]), "=r"(v[15])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (ci < cin && cb < cout) {
            float* dst = gw + ((long long)k * cin + ci) * cout + cb;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              red_add_v4(dst + 4 * q, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                         __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        ++k;
        while (k < kvol && s_first[k + 1] <= g) ++k;  // skip offsets without pairs
      }
    } else if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = make_idesc_tf32(CO) | (1u << 15) | (1u << 16);  // A, B MN-major
      int stage = 0;
      uint32_t phase = 0;
      int g = g_begin;
      while (g < g_end) {
        const int seg_end = min(g_end, s_first[k + 1]);
        uint32_t accumulate = 0;
        for (; g < seg_end; ++g) {
          mbar_wait(full_bar + stage, phase);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_smem = smem_u32(smem + stage * kWgStageBytes);
          const uint32_t b_smem = a_smem + a_bytes;
          for (int ks = 0; ks < WP / 8; ++ks) {
            const uint64_t a_desc = make_desc_mn_sw128(a_smem + ks * 2 * a_nb * 512, a_nb * 512);
            const uint64_t b_desc = make_desc_mn_sw128(b_smem + ks * 2 * b_nb * 512, b_nb * 512);
            umma_tf32(tmem_base, a_desc, b_desc, idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(empty_bar + stage);
          if (++stage == kWgStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(accum_bar);
        ++k;
        while (k < kvol && s_first[k + 1] <= g) ++k;
      }
    }
  }
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols)
                 : "memory");
  }
}

template <int CO>
int launch_wgrad_tc(const float* feat, const float* gout, const int* pairs, const int* num,
                    int64_t pair_stride, float* gw, int kvol, int cin, int cout, int inverse,
                    cudaStream_t stream) {
  DDF_SET_SMEM_ONCE(spconv_wgrad_tc_kernel<CO>, kWgSmemBytes);
  // persistent grid: 2 CTAs per SM (100 KB of smem each), fewer when the lists are short
  long long grid = 2 * ddf::kNumSM;
  const long long max_useful = ddf::cdiv((long long)pair_stride * kvol, 256);
  if (grid > max_useful) grid = max_useful;
  if (grid < 1) grid = 1;
  DDF_LAUNCH_PDL(spconv_wgrad_tc_kernel<CO>, (unsigned)grid, kThreadsTC, kWgSmemBytes, stream, feat, gout,
             pairs, num, (int)pair_stride, kvol, cin, cout, inverse, gw);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

}  // namespace

namespace ddf {

// true when the tensor-core kernel can take this launch
bool spconv_tc_supported(int kvol, int cin, int cout) {
  return kvol <= kMaxKvol && (cin % KCH == 0 || cin == 8 || cin == 16 || cin == 24) && cout >= 8 && cout <= 128;
}

// feat [n_in, cin]; wt [K, cout, cin] (K-major B operand); table [n_out, K]
int spconv_tc_launch(const float* feat, const float* wt, const int* table, const float* bias,
                     float* out, int64_t n_out, int kvol, int cin, int cout, cudaStream_t stream) {
  if (cout <= 16) return launch_tc<16>(feat, wt, table, bias, out, n_out, kvol, cin, cout, stream);
  if (cout <= 32) return launch_tc<32>(feat, wt, table, bias, out, n_out, kvol, cin, cout, stream);
  if (cout <= 64) return launch_tc<64>(feat, wt, table, bias, out, n_out, kvol, cin, cout, stream);
  return launch_tc<128>(feat, wt, table, bias, out, n_out, kvol, cin, cout, stream);
}


// kvol: the kernel keeps s_first[kvol + 1] in about 450 bytes of shared memory behind its barriers
bool spconv_wgrad_tc_supported(int kvol, int cin, int cout) {
  return kvol <= 100 && cin % 4 == 0 && cout % 16 == 0 && cin >= 4 && cin <= 128 && cout >= 16 && cout <= 128;
}

// gw must be zeroed by the caller; tiles are accumulated with red.global
int spconv_wgrad_tc_launch(const float* feat, const float* gout, const int* pairs, const int* num,
                           int64_t pair_stride, float* gw, int kvol, int cin, int cout, int inverse,
                           cudaStream_t stream) {
  if (cout <= 16) return launch_wgrad_tc<16>(feat, gout, pairs, num, pair_stride, gw, kvol, cin, cout, inverse, stream);
  if (cout <= 32) return launch_wgrad_tc<32>(feat, gout, pairs, num, pair_stride, gw, kvol, cin, cout, inverse, stream);
  if (cout <= 64) return launch_wgrad_tc<64>(feat, gout, pairs, num, pair_stride, gw, kvol, cin, cout, inverse, stream);
  return launch_wgrad_tc<128>(feat, gout, pairs, num, pair_stride, gw, kvol, cin, cout, inverse, stream);
}

}  // namespace ddf
