// C [M, N] = A^T . B for two tall row-major fp32 matrices A [K, M], B [K, N] (K = number of tokens, 1e5..1e6;
// M, N = layer widths): the weight gradient of every nn.Linear of the fusion encoder,
//   W.grad [out, in] = grad_out [T, out]^T . x [T, in]
// (<proj>/models/model_utils/actr_transformer.py:383-397 FFN, ops/modules/ms_deform_attn.py:124 value_proj, ...;
// autograd's AddmmBackward / MmBackward in the reference).  The library runs these "NT, both operands MN-major"
// products on Ampere-era cutlass_80 tf32 kernels at 160-170 TFLOP/s = 2.7 TB/s of the 598 MB hidden activation; the
// product is memory bound (0.5 flop per byte streamed at M = 128), so the job is to stream A and B once at the HBM
// rate.
//
// Design: split-K.  A CTA owns one (128-row, <= 256-column) tile of C as a TMEM accumulator and a contiguous range
// of the K dimension, which it streams in 32-row stages: (M_t + N_t) / 32 TMA boxes of [32 rows x 32 floats] per
// stage, SWIZZLE_128B_ATOM_32B, i.e. exactly the MN-major tf32 operand layout of tcgen05 (4-row groups of 512 bytes,
// 32-byte chunks XORed with the row; the layout csrc/sparse_conv_tma.cu builds by hand for the sparse wgrad); four
// kind::tf32 MMAs (K = 8 rows each) per stage; a ring of stages keeps >= 100 KB per SM in flight.  Warp 0 = TMA
// producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2..5 = epilogue: TMEM -> red.global.add.v4.f32 into the
// zeroed C (the split-K partial sums meet there).  Operands are read as tf32 by the hardware (top 19 bits), fp32
// accumulation: the arithmetic of the library's allow_tf32 path.
#include <cuda.h>

#include "common.cuh"

namespace {
constexpr int KS = 32;             // rows (K extent) per stage
constexpr int kBoxBytes = KS * 128;  // one [32 rows x 32 floats] box
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

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
      printf("xty mbarrier timeout: block %d thread %d smem 0x%x parity %u\n", blockIdx.x, threadIdx.x, addr, parity);
      __trap();
    }
  } while (!done);
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
// MN-major operand, SWIZZLE_128B_BASE32B: 4-row groups of 512 bytes (SBO), 32-channel blocks one TMA box apart (LBO)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(kBoxBytes >> 4) << 16;   // LBO: next 32-column block = next box
  d |= (uint64_t)(512 >> 4) << 32;         // SBO: next 4-row group
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                  // SWIZZLE_128B_BASE32B
  return d;
}
// tf32 x tf32 -> f32, M = 128, A and B MN-major
__device__ __forceinline__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
}

__global__ void __launch_bounds__(kThreads)
xty_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* __restrict__ c,
           int K, int M, int N, int n_tile, int m_tiles, int n_tiles, int splits, int n_stages) {
  ddf::pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tile = blockIdx.x % (m_tiles * n_tiles), split = blockIdx.x / (m_tiles * n_tiles);
  const int m0 = (tile / n_tiles) * 128, n0 = (tile % n_tiles) * n_tile;
  const int mt = min(128, M - m0), nt = min(n_tile, N - n0);   // multiples of 32
  const int mb = mt >> 5, nb = nt >> 5;
  const int stage_bytes = (4 + (n_tile >> 5)) * kBoxBytes;      // A always has room for 4 boxes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + n_stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = full + kMaxStages;
  uint64_t* accum_bar = empty + kMaxStages;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = (K + KS - 1) / KS;
  const int st_begin = (int)((long long)NS * split / splits), st_end = (int)((long long)NS * (split + 1) / splits);
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)nt) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && work) {
      int s = 0;
      uint32_t ph = 0;
      for (int st = st_begin; st < st_end; ++st) {
        mbar_wait(empty + s, ph ^ 1u);
        mbar_expect_tx(full + s, (uint32_t)((mb + nb) * kBoxBytes));
        const uint32_t a_dst = smem_u32(smem + s * stage_bytes), b_dst = a_dst + 4 * kBoxBytes;
        for (int i = 0; i < mb; ++i) tma_tile_2d(a_dst + i * kBoxBytes, &map_a, full + s, m0 + 32 * i, st * KS);
        for (int i = 0; i < nb; ++i) tma_tile_2d(b_dst + i * kBoxBytes, &map_b, full + s, n0 + 32 * i, st * KS);
        if (++s == n_stages) { s = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && work) {
      const uint32_t idesc = make_idesc(nt);
      int s = 0;
      uint32_t ph = 0;
      for (int st = st_begin; st < st_end; ++st) {
        mbar_wait(full + s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_smem = smem_u32(smem + s * stage_bytes), b_smem = a_smem + 4 * kBoxBytes;
#pragma unroll
        for (int ks = 0; ks < KS / 8; ++ks)   // 8 rows = two 4-row groups = 1024 bytes
          umma_tf32(tmem_base, make_desc_mn(a_smem + ks * 1024), make_desc_mn(b_smem + ks * 1024), idesc,
                    (st != st_begin || ks != 0) ? 1u : 0u);
        umma_commit(empty + s);
        if (++s == n_stages) { s = 0; ph ^= 1u; }
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  } else if (work) {
    // ===================== epilogue: accumulator -> C (red.global.add; split-K partial sums meet in C) ============
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    for (int cb = 0; cb < nt; cb += 16) {
      uint32_t v[16];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < mt) {
        float* dst = c + (long long)(m0 + row) * N + n0 + cb;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          red_add_v4(dst + 4 * i, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                     __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
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
// row-major fp32 [rows, cols]; box = [32 rows x 32 floats], 128-byte swizzle with 32-byte atoms, zero fill past the end
bool make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)KS};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

// Shapes the kernel takes: widths that are multiples of 32 (every Linear of the 3D-DF encoders), 16-byte aligned rows.
extern "C" int ddf_xty_supported(int64_t K, int64_t M, int64_t N) {
  return encode_fn() != nullptr && K > 0 && K < (1ll << 31) - 64 && M > 0 && N > 0 && M % 32 == 0 && N % 32 == 0 &&
         M <= 65536 && N <= 65536;
}

// c [M, N] = a [K, M]^T . b [K, N]   (fp32 row-major, tf32 products, fp32 accumulation; c is overwritten)
extern "C" int ddf_xty_tf32(const float* a, const float* b, float* c, int64_t K, int64_t M, int64_t N, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(ddf_xty_supported(K, M, N), "xty: unsupported shape K=%lld M=%lld N=%lld (widths must be multiples of 32)",
                (long long)K, (long long)M, (long long)N);
  DDF_CHECK_ARG(a && b && c, "xty: null pointer");
  DDF_CHECK_ARG(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0 && ((uintptr_t)c & 15) == 0, "xty: misaligned pointer");
  DDF_CUDA(cudaMemsetAsync(c, 0, sizeof(float) * (size_t)(M * N), stream));
  CUtensorMap map_a, map_b;
  if (!make_map(&map_a, a, K, M) || !make_map(&map_b, b, K, N)) {
    ddf::set_error("xty: cuTensorMapEncodeTiled failed (K=%lld M=%lld N=%lld)", (long long)K, (long long)M, (long long)N);
    return DDF_ERR_CUDA;
  }
  const int n_tile = N >= 256 ? 256 : (int)N;
  const int m_tiles = (int)ddf::cdiv(M, 128), n_tiles = (int)ddf::cdiv(N, n_tile);
  const int tiles = m_tiles * n_tiles;
  const long long NS = ddf::cdiv(K, KS);
  // one CTA per SM; every CTA streams at least 8 stages
  long long splits = ddf::kNumSM / tiles;
  if (splits < 1) splits = 1;
  if (splits > NS / 8) splits = NS / 8 > 0 ? NS / 8 : 1;
  const int stage_bytes = (4 + n_tile / 32) * kBoxBytes;
  int n_stages = (200 * 1024) / stage_bytes;
  if (n_stages > kMaxStages) n_stages = kMaxStages;
  const int smem = n_stages * stage_bytes + 256 + 1024;
  DDF_SET_SMEM_ONCE(xty_kernel, 227 * 1024);
  DDF_LAUNCH_PDL(xty_kernel, (unsigned)(tiles * splits), kThreads, smem, stream, map_a, map_b, c, (int)K, (int)M, (int)N,
             n_tile, m_tiles, n_tiles, (int)splits, n_stages);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}
