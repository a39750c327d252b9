// FFN of the 3D-DF encoder layers, forward, as ONE kernel:
//     h = dropout(relu(x W1^T + b1))      [T, F]   (kept for backward)
//     y = h W2^T + b2                     [T, D]
// (<proj>/models/model_utils/actr_transformer.py:383-397 forward_ffn: linear2(dropout(activation(linear1(src))));
// D = d_model = 128, F = d_ffn = 1024, T = 146k tokens at the TransFusion config.)
//
// The module chain was: library GEMM (writes 598 MB), one in-place bias / ReLU / dropout pass (reads + writes 598 MB),
// library GEMM (reads 598 MB) = 0.44 ms per FFN.  Here a CTA owns 128 tokens: the x tile is staged once by TMA, the
// hidden dimension is walked in chunks of 64: GEMM1 (tcgen05 kind::tf32, fp32 accumulators in TMEM, double buffered)
// -> epilogue warps (TMEM -> registers: bias, ReLU, counter-hash dropout; the chunk goes to shared memory in the
// K-major SWIZZLE_128B operand layout, from where a TMA store writes it to h - coalesced, instead of 32 scattered
// 16-byte stores per warp instruction) -> GEMM2 accumulates y in a second TMEM accumulator.  W1 / W2 chunks come through two 2-stage TMA rings.  The hidden activation is written once and never
// read again in forward; the kernel is bound by that write.
// Operands are read as tf32 by the tensor core (the arithmetic of the library's allow_tf32 path this replaces).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {
constexpr int TT = 128;            // tokens per CTA = UMMA M
constexpr int DM = 128;            // d_model
constexpr int FC = 64;             // hidden chunk
constexpr int kEpiWarps = 16;      // four per TMEM lane quadrant: each owns 16 columns of a hidden chunk.  The epilogue
                                   // (bias, ReLU, dropout hash, two stores per element) is as many lane-operations per
                                   // chunk as the SM issues in one GEMM1 + GEMM2 period: it needs every scheduler busy
constexpr int kThreads = (4 + kEpiWarps) * 32;   // warp 0: TMA loads, 1: GEMM1 issuer (+ TMEM), 2: GEMM2 issuer, 3: h store,
                                                 // warps 4..19: epilogue
constexpr int kXBytes = TT * DM * 4;         // 64 KB: 4 K-chunks [128 x 32 floats]
constexpr int kW1Bytes = FC * DM * 4;        // 32 KB: 4 K-chunks [64 x 32 floats]
constexpr int kW2Bytes = DM * FC * 4;        // 32 KB: 2 K-chunks [128 x 32 floats]
constexpr int kHsBytes = TT * FC * 4;        // 32 KB: 2 K-chunks [128 x 32 floats]
constexpr int kSmemBytes = kXBytes + 2 * kW1Bytes + 2 * kW2Bytes + kHsBytes + 256 + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done, spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 24)) {
      printf("ffn mbarrier timeout: block %d thread %d smem 0x%x parity %u\n", blockIdx.x, threadIdx.x, addr, parity);
      __trap();
    }
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_tile_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int col, int row) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(smem_u32(bar)), "r"(col), "r"(row)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
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
__device__ __forceinline__ constexpr uint32_t make_idesc_tf32(int n) {   // tf32 x tf32 -> f32, M = 128, K-major A and B
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// Dropout: counter-based, nothing stored (backward reads the pattern off h != 0).  The epilogue is instruction bound
// (16 warps x ~200 instructions per chunk against 1k cycles of tensor work: in-kernel trace), so the random bits are
// cheap: ONE round of the murmur3 finaliser over (4-element vector index, seed) gives four 8-bit uniforms, compared
// with the threshold by one SIMD instruction.  The drop probability is therefore quantised to 1 / 256
// (p = 0.1 -> 26 / 256); the kept values are scaled by the reciprocal of the quantised keep probability.
__device__ __forceinline__ uint32_t fmix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}
// bit j set = element j of vector i is kept; thr4 = the 8-bit threshold replicated in the four bytes
__device__ __forceinline__ unsigned keep4(uint32_t seed, uint32_t i, uint32_t thr4) {
  const uint32_t m = __vcmpgeu4(fmix32(i * 0x9E3779B1u + seed), thr4);     // 0xff per kept byte
  return (m & 1u) | ((m >> 7) & 2u) | ((m >> 14) & 4u) | ((m >> 21) & 8u);
}

// W1 [F, D] and W2 [D, F] re-laid as the shared-memory images of their 64-wide hidden chunks (K-major, SWIZZLE_128B:
// rows of 128 bytes, 16-byte pieces XORed with (row & 7)), so that a chunk arrives by ONE linear bulk copy of 32 KB
// instead of 256 tensor-map rows of 128 bytes: the TMA unit needs 4-5 cycles per such row and re-streaming 1 MB of
// weights per 128-token tile that way was 2/3 of the kernel's time.
//   p1 [F / 64][4 K-chunks][64 rows][32 floats],  p2 [F / 64][2 K-chunks][128 rows][32 floats]
__global__ void __launch_bounds__(256)
ffn_pack_weights_kernel(const float* __restrict__ w1, const float* __restrict__ w2, float* __restrict__ p1,
                        float* __restrict__ p2, int F) {
  ddf::pdl_sync();
  const int t = blockIdx.x * 256 + threadIdx.x;          // one 16-byte piece each
  const int n1 = F * DM / 4;
  if (t < n1) {
    const int ch = t & 7, r = (t >> 3) & 63, kc = (t >> 9) & 3, c = t >> 11;
    const float4 v = ldg4(w1 + (long long)(c * FC + r) * DM + kc * 32 + ch * 4);
    *reinterpret_cast<float4*>(p1 + ((((long long)c * 4 + kc) * 64 + r) * 32) + ((ch ^ (r & 7)) << 2)) = v;
  } else if (t < 2 * n1) {
    const int u = t - n1;
    const int ch = u & 7, r = (u >> 3) & 127, kc = (u >> 10) & 1, c = u >> 11;
    const float4 v = ldg4(w2 + (long long)r * F + c * FC + kc * 32 + ch * 4);
    *reinterpret_cast<float4*>(p2 + ((((long long)c * 2 + kc) * 128 + r) * 32) + ((ch ^ (r & 7)) << 2)) = v;
  }
}

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
ffn_fwd_kernel(const __grid_constant__ CUtensorMap map_x, const float* __restrict__ p1,
               const float* __restrict__ p2, const __grid_constant__ CUtensorMap map_h,
               const float* __restrict__ b1, const float* __restrict__ b2, float* __restrict__ y, int T, int F,
               unsigned long long seed, unsigned thr, float scale) {
  ddf::pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* xs = smem;
  uint8_t* w1s = xs + kXBytes;
  uint8_t* w2s = w1s + 2 * kW1Bytes;
  uint8_t* hs = w2s + 2 * kW2Bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(hs + kHsBytes);
  uint64_t* x_full = bars;            // 1
  uint64_t* w1_full = bars + 1;       // 2
  uint64_t* w1_empty = bars + 3;      // 2
  uint64_t* w2_full = bars + 5;       // 2
  uint64_t* w2_empty = bars + 7;      // 2
  uint64_t* acc1_full = bars + 9;     // 2
  uint64_t* acc1_free = bars + 11;    // 2
  uint64_t* hs_full = bars + 13;      // 2: one per 32-column half of Hs (= K-chunk of GEMM2)
  uint64_t* hs_empty = bars + 15;     // 4: [half][chunk parity] (a waiter must see every phase of its barrier)
  uint64_t* acc2_full = bars + 19;    // 1
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 20);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * TT;
  const int NC = F / FC;
  constexpr uint32_t kTmemCols = 256;          // acc1: 2 x 64 columns, acc2: 128 columns
  constexpr int kEpiThreads = kEpiWarps * 32;
#ifdef DDF_TRACE
  __shared__ long long tr[10][16];
  const bool trb = blockIdx.x == 200;
#define FTR(e, c) do { if (trb && (c) < 16) tr[e][c] = clock64(); } while (0)
#else
#define FTR(e, c) do {} while (0)
#endif

  if (tid == 0) {
    mbar_init(x_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(w1_full + s, 1);
      mbar_init(w1_empty + s, 1);
      mbar_init(w2_full + s, 1);
      mbar_init(w2_empty + s, 1);
      mbar_init(acc1_full + s, 1);
      mbar_init(acc1_free + s, kEpiThreads / 2);
    }
    for (int i = 0; i < 2; ++i) mbar_init(hs_full + i, kEpiThreads / 4);   // the 4 warps that own this half
    for (int i = 0; i < 4; ++i) mbar_init(hs_empty + i, 2);   // GEMM2 has read the half AND the TMA store of h has read it
    mbar_init(acc2_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  const uint32_t acc2 = tmem_base + 128;

  if (warp == 0) {
    // ===================== TMA loads: the x tile, then the W1 / W2 chunk rings =====================
    if (lane == 0) {
      mbar_expect_tx(x_full, kXBytes);
      for (int kc = 0; kc < DM / 32; ++kc) tma_tile_2d(smem_u32(xs + kc * (TT * 128)), &map_x, x_full, kc * 32, row0);
      // the two rings advance independently: W1 chunk c + 2 can be fetched as soon as GEMM1(c) is done, long before
      // GEMM2(c) releases the W2 stage
      int c1 = 0, c2 = 0;
      while (c1 < NC || c2 < NC) {
        if (c1 < NC && (c2 >= NC || c1 <= c2 + 1)) {
          const int s = c1 & 1;
          mbar_wait(w1_empty + s, ((uint32_t)(c1 >> 1) & 1u) ^ 1u);
          FTR(0, c1);
          mbar_expect_tx(w1_full + s, kW1Bytes);
          bulk_load(smem_u32(w1s + s * kW1Bytes), p1 + (long long)c1 * (kW1Bytes / 4), kW1Bytes, w1_full + s);
          ++c1;
        } else {
          const int s = c2 & 1;
          mbar_wait(w2_empty + s, ((uint32_t)(c2 >> 1) & 1u) ^ 1u);
          FTR(1, c2);
          mbar_expect_tx(w2_full + s, kW2Bytes);
          bulk_load(smem_u32(w2s + s * kW2Bytes), p2 + (long long)c2 * (kW2Bytes / 4), kW2Bytes, w2_full + s);
          ++c2;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== GEMM1 issuer: acc1[c & 1] = x . W1_c^T =====================
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc_tf32(FC);
      const uint64_t a0 = make_desc_sw128(smem_u32(xs));
      const uint64_t b0 = make_desc_sw128(smem_u32(w1s));
      mbar_wait(x_full, 0);
      for (int c = 0; c < NC; ++c) {
        const int s = c & 1;
        if (c >= 2) mbar_wait(acc1_free + s, (uint32_t)((c >> 1) - 1) & 1u);   // the epilogue has read chunk c - 2
        mbar_wait(w1_full + s, (uint32_t)(c >> 1) & 1u);
        FTR(2, c);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t bs = b0 + (uint64_t)(s * (kW1Bytes >> 4));
#pragma unroll
        for (int kc = 0; kc < DM / 32; ++kc) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)       // descriptor address fields are in 16-byte units
            umma_tf32(tmem_base + (uint32_t)(s * FC), a0 + (uint64_t)(kc * ((TT * 128) >> 4) + 2 * ks),
                      bs + (uint64_t)(kc * ((FC * 128) >> 4) + 2 * ks), idesc1, (kc | ks) ? 1u : 0u);
        }
        umma_commit(w1_empty + s);
        umma_commit(acc1_full + s);
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ===================== GEMM2 issuer: y += Hs_c . W2_c^T =====================
    if (lane == 0) {
      constexpr uint32_t idesc2 = make_idesc_tf32(DM);
      const uint64_t a0 = make_desc_sw128(smem_u32(hs));
      const uint64_t b0 = make_desc_sw128(smem_u32(w2s));
      for (int c = 0; c < NC; ++c) {
        const int s = c & 1;
        mbar_wait(w2_full + s, (uint32_t)(c >> 1) & 1u);
        const uint64_t bs = b0 + (uint64_t)(s * (kW2Bytes >> 4));
#pragma unroll
        for (int kc = 0; kc < FC / 32; ++kc) {
          // the halves of Hs are handed over separately: the 4 MMAs of one half run while the other is being written
          mbar_wait(hs_full + kc, (uint32_t)c & 1u);
          if (kc == 0) FTR(3, c);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32(acc2, a0 + (uint64_t)(kc * ((TT * 128) >> 4) + 2 * ks),
                      bs + (uint64_t)(kc * ((DM * 128) >> 4) + 2 * ks), idesc2, (c | kc | ks) ? 1u : 0u);
          umma_commit(hs_empty + kc * 2 + s);
        }
        umma_commit(w2_empty + s);
        FTR(5, c);
      }
      umma_commit(acc2_full);
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================== h store: each half of Hs IS the swizzled image of one TMA box =====================
    if (lane < 2) {
      const int kc = lane;                         // lane = half
      for (int c = 0; c < NC; ++c) {
        mbar_wait(hs_full + kc, (uint32_t)c & 1u);  // written and fenced (fence.proxy.async) by the epilogue warps
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_h),
                     "r"(smem_u32(hs + kc * (TT * 128))), "r"(c * FC + kc * 32), "r"(row0)
                     : "memory");                   // rows past T are clipped by the tensor map
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (kc == 0) FTR(4, c);
        mbar_arrive(hs_empty + kc * 2 + (c & 1));
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // h is in global memory before the CTA ends
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    // Two groups of 8 warps ping-pong over the chunks (group = chunk parity, the parity of the acc1 buffer too): while
    // one group waits for Hs / writes it, the other one computes.  thread = (token row, 32 columns of the chunk).
    const int q = warp & 3;                       // TMEM lane quadrant of this warp
    const int grp = (warp - 4) >> 3;              // chunk parity this warp serves
    const int half = ((warp - 4) >> 2) & 1;       // which 32 columns of a hidden chunk = which K-chunk of Hs
    const int r = q * 32 + lane;                  // row of the tile
    const long long grow = (long long)row0 + r;
    const bool live = grow < T;
    const bool t0 = warp == 4 && lane == 0;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(grp * FC + half * 32);
    // K-major rows of 128 bytes per 32-column chunk, 16-byte pieces XORed with (row & 7)
    const uint32_t hs_row = smem_u32(hs) + (uint32_t)(half * (TT * 128) + r * 128);
    const int rx = r & 7;
    const uint32_t vbase = (uint32_t)((grow * F) >> 2) + (uint32_t)(half * 8);   // 4-element vector index of (row, col)
    const uint32_t seed32 = (uint32_t)seed ^ (uint32_t)(seed >> 32);
    const float* bias = b1 + half * 32;
    for (int c = grp; c < NC; c += 2) {
      if (t0) FTR(6, c >> 1);
      float4 bv[8];                              // this thread's 32 bias values: in flight while it waits for GEMM1
#pragma unroll
      for (int j = 0; j < 8; ++j) bv[j] = ldg4(bias + c * FC + 4 * j);
      mbar_wait(acc1_full + grp, (uint32_t)(c >> 1) & 1u);
      if (t0) FTR(7, c >> 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v[32];
      {
        uint32_t lo[16], hi[16];
        tmem_ld16(lane_addr, lo);
        tmem_ld16(lane_addr + 16, hi);
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i] = lo[i]; v[16 + i] = hi[i]; }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(acc1_free + grp);              // GEMM1(c + 2) may overwrite this accumulator
      float o[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const unsigned k = thr ? keep4(seed32, vbase + (uint32_t)(c * (FC / 4) + j), thr) : 15u;
        const float m0 = (k & 1u) ? scale : 0.f, m1 = (k & 2u) ? scale : 0.f;
        const float m2 = (k & 4u) ? scale : 0.f, m3 = (k & 8u) ? scale : 0.f;
        o[4 * j] = fmaxf(__uint_as_float(v[4 * j]) + bv[j].x, 0.f) * m0;
        o[4 * j + 1] = fmaxf(__uint_as_float(v[4 * j + 1]) + bv[j].y, 0.f) * m1;
        o[4 * j + 2] = fmaxf(__uint_as_float(v[4 * j + 2]) + bv[j].z, 0.f) * m2;
        o[4 * j + 3] = fmaxf(__uint_as_float(v[4 * j + 3]) + bv[j].w, 0.f) * m3;
      }
      if (t0) FTR(8, c >> 1);
      // GEMM2(c - 1) and the store of chunk c - 1 (the other group's chunk) have read Hs
      if (c > 0) mbar_wait(hs_empty + half * 2 + (grp ^ 1), (uint32_t)((c - 1) >> 1) & 1u);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hs_row + (uint32_t)((j ^ rx) << 4)), "f"(o[4 * j]),
                     "f"(o[4 * j + 1]), "f"(o[4 * j + 2]), "f"(o[4 * j + 3])
                     : "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(hs_full + half);
      if (t0) FTR(9, c >> 1);
    }
    // y tile: this thread's 32 columns
    mbar_wait(acc2_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int g = 0; g < DM / 64; ++g) {
      uint32_t v[16];
      const int c0 = (grp * 2 + half) * (DM / 4) + g * 16;
      tmem_ld16(acc2 + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (live) {
        float* dst = y + grow * DM + c0;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b = b2 ? ldg4(b2 + c0 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(dst + i) =
              make_float4(__uint_as_float(v[i]) + b.x, __uint_as_float(v[i + 1]) + b.y, __uint_as_float(v[i + 2]) + b.z,
                          __uint_as_float(v[i + 3]) + b.w);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
#ifdef DDF_TRACE
  if (trb && tid == 0) {
    const long long t0 = tr[2][0];
    for (int c = 0; c < 16 && c < NC; ++c)
      printf("ffn %2d: W1 slot %6lld W2 slot %6lld | g1 w1_full %6lld | g2 saw hs_full %6lld store read done %6lld g2 issued %6lld | epi(even chunks) start %6lld acc1_full %6lld computed %6lld wrote %6lld\n",
             c, tr[0][c] - t0, tr[1][c] - t0, tr[2][c] - t0, tr[3][c] - t0, tr[4][c] - t0, tr[5][c] - t0, tr[6][c] - t0,
             tr[7][c] - t0, tr[8][c] - t0, tr[9][c] - t0);
  }
#endif
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// row-major fp32 [rows, cols]; box = [box_rows x 32 floats], SWIZZLE_128B (K-major operand tiles), zero fill past the end
bool make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
}  // namespace

// Shapes the fused kernel takes: d_model 128, d_ffn a multiple of 64.
extern "C" int ddf_ffn_supported(int64_t T, int64_t D, int64_t F) {
  return encode_fn() != nullptr && T > 0 && T < (1ll << 31) - 256 && D == DM && F >= FC && F % FC == 0 && F <= 65536;
}

// h [T, F] = dropout(relu(x [T, D] . w1 [F, D]^T + b1)), y [T, D] = h . w2 [D, F]^T + b2.  fp32 row-major, tf32 products,
// fp32 accumulation; p = drop probability (0 in eval), kept values scaled by 1 / (1 - p); the keep decisions are a
// counter hash of (seed, element index); ddf_bias_relu_dropout_backward applies to h unchanged (it reads h != 0).
extern "C" int64_t ddf_ffn_workspace_bytes(int64_t D, int64_t F) { return D > 0 && F > 0 ? 2 * D * F * 4 : -1; }

// 8-bit drop threshold of the kernel: u < t8 is dropped, t8 = round(256 p), at least 1 when p > 0
static unsigned drop_threshold8(float p) {
  if (!(p > 0.f)) return 0;
  unsigned t8 = (unsigned)(p * 256.f + 0.5f);
  return t8 < 1 ? 1 : (t8 > 255 ? 255 : t8);
}
// The drop probability ddf_ffn_forward really applies for a requested p (quantised to 1 / 256): backward must scale
// by 1 / (1 - this), not by 1 / (1 - p).
extern "C" float ddf_ffn_dropout_p(float p) { return (float)drop_threshold8(p) / 256.f; }

extern "C" int ddf_ffn_forward(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                               float* h, float* y, void* workspace, int64_t T, int64_t D, int64_t F, float p,
                               uint64_t seed, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(T >= 0 && p >= 0.f && p < 1.f, "ffn_forward: bad arguments");
  if (T == 0) return DDF_OK;
  DDF_CHECK_ARG(ddf_ffn_supported(T, D, F), "ffn_forward: unsupported shape T=%lld D=%lld F=%lld (D = 128, F %% 64 == 0)",
                (long long)T, (long long)D, (long long)F);
  DDF_CHECK_ARG(x && w1 && b1 && w2 && h && y && workspace, "ffn_forward: null pointer");
  DDF_CHECK_ARG(aligned16(workspace), "ffn_forward: misaligned workspace");
  DDF_CHECK_ARG(aligned16(x) && aligned16(w1) && aligned16(b1) && aligned16(w2) && aligned16(b2) && aligned16(h) && aligned16(y),
                "ffn_forward: misaligned pointer");
  CUtensorMap map_x, map_h;
  if (!make_map(&map_x, x, T, D, TT) || !make_map(&map_h, h, T, F, TT)) {
    ddf::set_error("ffn_forward: cuTensorMapEncodeTiled failed");
    return DDF_ERR_CUDA;
  }
  // drop threshold on 8-bit uniforms: u < t8 is dropped, t8 = round(256 p) (at least 1 when p > 0)
  unsigned thr = 0;
  float scale = 1.f;
  if (p > 0.f) {
    const unsigned t8 = drop_threshold8(p);
    thr = t8 * 0x01010101u;
    scale = 256.f / (float)(256 - t8);
  }
  float* p1 = reinterpret_cast<float*>(workspace);
  float* p2 = p1 + D * F;
  DDF_LAUNCH(ffn_pack_weights_kernel, (unsigned)ddf::cdiv(2 * F * D / 4, 256), 256, 0, stream, w1, w2, p1, p2, (int)F);
  DDF_SET_SMEM_ONCE(ffn_fwd_kernel, kSmemBytes);
  DDF_LAUNCH_PDL(ffn_fwd_kernel, (unsigned)ddf::cdiv(T, TT), kThreads, kSmemBytes, stream, map_x, p1, p2, map_h, b1, b2, y,
             (int)T, (int)F, (unsigned long long)seed, thr, scale);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}
