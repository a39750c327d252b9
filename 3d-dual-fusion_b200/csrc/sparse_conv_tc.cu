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
  const int n_chunks = cin / KCH;

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
          const int j = s_idx[r * kvol + k];
          const float* src = feat + (j >= 0 ? ((long long)j * cin + c0 + ch * 4) : 0);
          cp_async16(a_smem + r * 128 + ((ch ^ (r & 7)) << 4), src, j >= 0 ? 16u : 0u);
        }
#pragma unroll
        for (int i = 0; i < (CO * 8 + kProducers - 1) / kProducers; ++i) {
          const int e = i * kProducers + tid;
          if (e < CO * 8) {
            const int r = e >> 3, ch = e & 7;
            const bool ok = r < cout;
            const float* src = wk + (ok ? ((long long)r * cin + c0 + ch * 4) : 0);
            cp_async16(b_smem + r * 128 + ((ch ^ (r & 7)) << 4), src, ok ? 16u : 0u);
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
#pragma unroll
          for (int ks = 0; ks < KCH / 8; ++ks) {
            // advance 8 tf32 = 32 bytes along K inside the 128-byte swizzle row: +2 in 16-byte units
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
  static bool configured = false;
  if (!configured) {
    DDF_CUDA(cudaFuncSetAttribute(spconv_tc_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::kSmemBytes));
    configured = true;
  }
  const unsigned grid = (unsigned)ddf::cdiv(n_out, TM);
  DDF_LAUNCH(spconv_tc_kernel<CO>, grid, kThreadsTC, Cfg::kSmemBytes, stream, feat, wt, table, bias,
             out, (int)n_out, kvol, cin, cout);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}


// ------------------------------------------------------------------------------------------------
// wgrad on tensor cores:  gW[k] (Cin x Cout) = sum over the pairs (i, o) of offset k of
//   feat[i, :]^T . gout[o, :]
// = a GEMM with M = Cin (padded to 128 TMEM lanes), N = Cout, K = pairs.  The gathered rows are
// channel-contiguous, i.e. both operands are MN-major.  For 32-bit MN-major operands the only UMMA
// smem layout is SWIZZLE_128B_BASE32B: atoms of 4 pairs x 32 channels (4 rows of 128 bytes), the
// 32-byte chunk index XORed with (pair & 3).  Atoms of one 4-pair group are contiguous (LBO = 512 B
// between 32-channel blocks, SBO between 4-pair groups); one tf32 MMA (K = 8) spans two groups.
// Grid (kvol, S): CTA (k, s) reduces slice s of pair list k in stages of 32 pairs (4 MMAs, K = 8)
// into TMEM and adds its tile to gW[k] with 16-byte red.global.
// ------------------------------------------------------------------------------------------------
constexpr int WPAIRS = 32;  // pairs per stage

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

template <int CO>
struct WgCfg {
  static constexpr int kStages = 3;
  static constexpr int kABytes = WPAIRS * 128 * 4;   // M padded to 128 channels: 16 KB
  static constexpr int kBBytes = WPAIRS * CO * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 256 + 1024;
};

template <int CO>
__global__ void __launch_bounds__(kThreadsTC)
spconv_wgrad_tc_kernel(const float* __restrict__ feat, const float* __restrict__ gout,
                       const int* __restrict__ pairs, const int* __restrict__ num, int pair_stride,
                       int cin, int cout, int inverse, float* __restrict__ gw) {
  using Cfg = WgCfg<CO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* accum_bar = bars + 2 * Cfg::kStages;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k = blockIdx.x;
  const int nk = num[k];
  const int S = gridDim.y;
  // slices are multiples of the stage size so only the last stage of a slice is ragged
  int per = (nk + S - 1) / S;
  per = ((per + WPAIRS - 1) / WPAIRS) * WPAIRS;
  const int s0 = blockIdx.y * per;
  const int s1 = min(nk, s0 + per);
  if (s0 >= s1) return;  // uniform for the CTA, before any barrier / TMEM allocation
  const int n_stage = (s1 - s0 + WPAIRS - 1) / WPAIRS;
  const int* pin = pairs + ((long long)k * 2 + (inverse ? 1 : 0)) * pair_stride;
  const int* pout = pairs + ((long long)k * 2 + (inverse ? 0 : 1)) * pair_stride;
  constexpr int kTmemCols = CO < 32 ? 32 : CO;

  if (tid == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar + s, kProducers);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // channel blocks of A beyond cin are zero for the whole kernel: clear them once
  const int a_blocks = cin / 32;  // live 32-channel blocks of the 4
  if (a_blocks < 4) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      uint8_t* a = smem + s * Cfg::kStageBytes;
      for (int e = tid; e < Cfg::kABytes / 16; e += kThreadsTC) {
        const int atom = e >> 5;  // 32 chunks of 16 B per 512-B atom
        if ((atom & 3) >= a_blocks) reinterpret_cast<float4*>(a)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  const int nb = CO / 32;           // 32-channel blocks of B (CO is the padded width)
  const int b_live = cout / 32;

  if (warp < 4) {
    int stage = 0;
    uint32_t phase = 0;
    const int a_chunks = a_blocks * 8;  // 16-B chunks per pair row of A
    const int b_chunks = b_live * 8;
    for (int it = 0; it < n_stage; ++it) {
      mbar_wait(empty_bar + stage, phase ^ 1u);
      const uint32_t a_smem = smem_u32(smem + stage * Cfg::kStageBytes);
      const uint32_t b_smem = a_smem + Cfg::kABytes;
      const int p0 = s0 + it * WPAIRS;
      for (int e = tid; e < WPAIRS * a_chunks; e += kProducers) {
        const int p = e / a_chunks, cc = e % a_chunks;
        const int mb = cc >> 3, ch = cc & 7;
        const bool ok = p0 + p < s1;
        const int row = ok ? __ldg(pin + p0 + p) : 0;
        const uint32_t dst = a_smem + mn_chunk_offset(p, mb, ch, 4);
        cp_async16(dst, feat + (long long)row * cin + cc * 4, ok ? 16u : 0u);
      }
      for (int e = tid; e < WPAIRS * b_chunks; e += kProducers) {
        const int p = e / b_chunks, cc = e % b_chunks;
        const int mb = cc >> 3, ch = cc & 7;
        const bool ok = p0 + p < s1;
        const int row = ok ? __ldg(pout + p0 + p) : 0;
        const uint32_t dst = b_smem + mn_chunk_offset(p, mb, ch, nb);
        cp_async16(dst, gout + (long long)row * cout + cc * 4, ok ? 16u : 0u);
      }
      cp_async_arrive_noinc(full_bar + stage);
      if (++stage == Cfg::kStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    // epilogue: lane = input channel, columns = output channels
    mbar_wait(accum_bar, 0);
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
  } else {
    if (lane == 0) {
      // M = 128, N = CO, both operands MN-major
      constexpr uint32_t idesc = make_idesc_tf32(CO) | (1u << 15) | (1u << 16);
      int stage = 0;
      uint32_t phase = 0, accumulate = 0;
      for (int it = 0; it < n_stage; ++it) {
        mbar_wait(full_bar + stage, phase);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_smem = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t b_smem = a_smem + Cfg::kABytes;
#pragma unroll
        for (int ks = 0; ks < WPAIRS / 8; ++ks) {
          const uint64_t a_desc = make_desc_mn_sw128(a_smem + ks * 2 * 4 * 512, 4 * 512);
          const uint64_t b_desc = make_desc_mn_sw128(b_smem + ks * 2 * nb * 512, nb * 512);
          umma_tf32(tmem_base, a_desc, b_desc, idesc, accumulate);
          accumulate = 1;
        }
        umma_commit(empty_bar + stage);
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
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
  using Cfg = WgCfg<CO>;
  static bool configured = false;
  if (!configured) {
    DDF_CUDA(cudaFuncSetAttribute(spconv_wgrad_tc_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::kSmemBytes));
    configured = true;
  }
  // ~4 waves of (offset, slice) CTAs, each slice at least 256 pairs when the list is that long
  int S = (int)ddf::cdiv(4 * ddf::kNumSM, kvol);
  const int maxS = (int)ddf::cdiv(pair_stride, 256);
  if (S > maxS) S = maxS;
  if (S < 1) S = 1;
  dim3 grid((unsigned)kvol, (unsigned)S);
  DDF_LAUNCH(spconv_wgrad_tc_kernel<CO>, grid, kThreadsTC, Cfg::kSmemBytes, stream, feat, gout, pairs,
             num, (int)pair_stride, cin, cout, inverse, gw);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

}  // namespace

namespace ddf {

// true when the tensor-core kernel can take this launch
bool spconv_tc_supported(int kvol, int cin, int cout) {
  return kvol <= kMaxKvol && cin % KCH == 0 && cin >= KCH && cout >= 8 && cout <= 128;
}

// feat [n_in, cin]; wt [K, cout, cin] (K-major B operand); table [n_out, K]
int spconv_tc_launch(const float* feat, const float* wt, const int* table, const float* bias,
                     float* out, int64_t n_out, int kvol, int cin, int cout, cudaStream_t stream) {
  if (cout <= 16) return launch_tc<16>(feat, wt, table, bias, out, n_out, kvol, cin, cout, stream);
  if (cout <= 32) return launch_tc<32>(feat, wt, table, bias, out, n_out, kvol, cin, cout, stream);
  if (cout <= 64) return launch_tc<64>(feat, wt, table, bias, out, n_out, kvol, cin, cout, stream);
  return launch_tc<128>(feat, wt, table, bias, out, n_out, kvol, cin, cout, stream);
}


bool spconv_wgrad_tc_supported(int cin, int cout) {
  return cin % 32 == 0 && cout % 32 == 0 && cin >= 32 && cin <= 128 && cout >= 32 && cout <= 128;
}

// gw must be zeroed by the caller; tiles are accumulated with red.global
int spconv_wgrad_tc_launch(const float* feat, const float* gout, const int* pairs, const int* num,
                           int64_t pair_stride, float* gw, int kvol, int cin, int cout, int inverse,
                           cudaStream_t stream) {
  if (cout <= 32) return launch_wgrad_tc<32>(feat, gout, pairs, num, pair_stride, gw, kvol, cin, cout, inverse, stream);
  if (cout <= 64) return launch_wgrad_tc<64>(feat, gout, pairs, num, pair_stride, gw, kvol, cin, cout, inverse, stream);
  return launch_wgrad_tc<128>(feat, gout, pairs, num, pair_stride, gw, kvol, cin, cout, inverse, stream);
}

}  // namespace ddf
