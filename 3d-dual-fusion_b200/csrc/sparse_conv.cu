// Sparse 3-D convolution (SubM / regular / inverse) forward, dgrad, wgrad for sm_100a — fp32 SIMT
// implicit-GEMM kernels (the tcgen05 tf32 path for wide layers lives in sparse_conv_tc.cu).
//
// Replaces indiceConv / indiceConvBackward
//   <TF>/ops/spconv/include/spconv/spconv_ops.h:260-361, 363-456
// which run, per kernel offset, gather kernel -> cuBLAS mm_out -> scatter-add kernel (<=27 x 3
// launches, every intermediate through HBM, plus an indiceNum D2H sync).  Here one launch per
// conv: an output-stationary kernel walks the gather table G[N_out, K] built with the rulebook
// (rulebook.cu), stages the gathered input rows and the filter slice of each offset in shared
// memory and keeps the [rows x Cout] accumulator tile in registers; the output is written once.
//   forward : out[o,:]  = sum_k in[G[o,k],:]   . W[k]        (W[k]: Cin x Cout)
//   dgrad   : gin[i,:]  = sum_k gout[GT[i,k],:]. W[k]^T      (same kernel, transposed filters)
//   wgrad   : gW[k]     = sum_{(i,o) in pairs[k]} in[i,:]^T gout[o,:]   (walks the pair lists)
// The SubM centre tap needs no special case here (the reference treats arg-max offset as identity,
// spconv_ops.h:271-303): its gather-table column is simply the identity.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace ddf {
// tcgen05 path (sparse_conv_tc.cu)
bool spconv_tc_supported(int kvol, int cin, int cout);
int spconv_tc_launch(const float* feat, const float* wt, const int* table, const float* bias,
                     float* out, int64_t n_out, int kvol, int cin, int cout, cudaStream_t stream);
// TMA-staged tcgen05 path (sparse_conv_tma.cu)
bool spconv_tma_supported(int kvol, int cin, int cout);
int spconv_tma_launch(const float* feat, const float* wt, const int* table, const float* bias, float* out,
                      int64_t n_out, int64_t n_in, int kvol, int cin, int cout, bool gather4, bool split,
                      cudaStream_t stream);
bool spconv_wgrad_table_supported(int kvol, int cin, int cout);
int spconv_wgrad_table_launch(const float* feat, const float* gout, const int* table, float* gw,
                              int64_t n_out, int64_t n_in, int kvol, int cin, int cout, cudaStream_t stream);
bool spconv_wgrad_tc_supported(int kvol, int cin, int cout);
int spconv_wgrad_tc_launch(const float* feat, const float* gout, const int* pairs, const int* num,
                           int64_t pair_stride, float* gw, int kvol, int cin, int cout, int inverse,
                           cudaStream_t stream);
}  // namespace ddf

namespace {

// Conv kernel selection. 0 (DDF_DISABLE_TC=1 or ddf_set_tensor_cores(0)): fp32 SIMT kernels, full
// fp32 products - A/B timing and the strict 1e-3 whole-path parity test. 1 (default): tcgen05 tf32,
// multi-tile kernel (filter slices by tiled TMA and shared by up to 4 row tiles, rows gathered by
// cp.async from 16 warps) where the layer shape allows, else the single-tile cp.async kernel.
// 2: tcgen05 tf32, single-tile cp.async kernel only. 3: as 1 with the rows gathered by TMA gather4.
// 4 (default): as 1, but forward and dgrad of the layers the multi-tile kernel takes run "bf16x3": both operands
// split into bf16 hi + lo halves, three MMAs per product pair (hi.hi + hi.lo + lo.hi), i.e. a 16-bit significand
// per product instead of tf32's 11 bits - this is what keeps the whole 21-conv path inside 1e-3 of the fp32
// reference (single-pass tf32 accumulates to about 1.8e-3 through the batch-statistics BatchNorms). wgrad stays tf32.
int g_tc_state = -1;  // -1: not read yet
int tc_mode_state() {
  if (g_tc_state < 0) {
    const char* e = getenv("DDF_DISABLE_TC");
    const char* m = getenv("DDF_TC_MODE");
    g_tc_state = (e && e[0] == '1') ? 0 : (m && m[0] >= '0' && m[0] <= '4') ? m[0] - '0' : 4;
  }
  return g_tc_state;
}
bool tc_enabled() { return tc_mode_state() != 0; }
// the warp-per-row fp32 kernel for narrow, low-density layers (every mode except 2, the A/B baseline)
bool sparse_rows_enabled() { return tc_mode_state() != 2; }
bool tma_enabled() { return tc_mode_state() == 1 || tc_mode_state() >= 3; }
bool bf16x3_enabled() { return tc_mode_state() == 4; }
bool gather4_enabled() { return tc_mode_state() == 3; }


constexpr int kThreads = 256;
constexpr int TM = 128;  // output rows per CTA
constexpr int KC = 16;   // input channels per smem chunk

// ------------------------------------------------------------------------------------------------
// forward / dgrad kernel.  CO = padded output-channel tile (16/32/64/128), thread tile RM x 4.
// ------------------------------------------------------------------------------------------------
template <int CO>
__global__ void __launch_bounds__(kThreads)
spconv_gather_gemm_kernel(const float* __restrict__ feat, const float* __restrict__ filt,
                          const int* __restrict__ table, const float* __restrict__ bias,
                          float* __restrict__ out, int n_out, int kvol, int cin, int cout,
                          int co_base) {
  ddf::pdl_sync();
  constexpr int TX = CO / 4;          // threads along channels
  constexpr int TY = kThreads / TX;   // threads along rows
  constexpr int RM = TM / TY;         // rows per thread
  __shared__ float sA[TM][KC + 1];
  __shared__ __align__(16) float sB[KC][CO];
  __shared__ int sIdx[TM];
  __shared__ int sAny;

  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int row0 = blockIdx.x * TM;
  float acc[RM][4];
#pragma unroll
  for (int r = 0; r < RM; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

  for (int k = 0; k < kvol; ++k) {
    __syncthreads();  // previous offset's sA/sB/sIdx fully consumed
    if (threadIdx.x == 0) sAny = 0;
    __syncthreads();
    if (threadIdx.x < TM) {
      const int o = row0 + threadIdx.x;
      const int j = o < n_out ? table[(long long)o * kvol + k] : -1;
      sIdx[threadIdx.x] = j;
      if (j >= 0) sAny = 1;
    }
    __syncthreads();
    if (!sAny) continue;  // no row of this tile has a neighbour at offset k
    const float* wk = filt + (long long)k * cin * cout;
    for (int c0 = 0; c0 < cin; c0 += KC) {
      if (c0) __syncthreads();
      // stage A: TM x KC gathered features (zeros for missing neighbours / channel tail)
      for (int e = threadIdx.x; e < TM * KC; e += kThreads) {
        const int r = e / KC, c = e % KC;
        const int j = sIdx[r];
        sA[r][c] = (j >= 0 && c0 + c < cin) ? __ldg(feat + (long long)j * cin + c0 + c) : 0.f;
      }
      // stage B: KC x CO filter slice
      for (int e = threadIdx.x; e < KC * CO; e += kThreads) {
        const int c = e / CO, n = e % CO;
        sB[c][n] = (c0 + c < cin && co_base + n < cout)
                       ? __ldg(wk + (long long)(c0 + c) * cout + co_base + n)
                       : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < KC; ++c) {
        const float4 b = *reinterpret_cast<const float4*>(&sB[c][tx * 4]);
#pragma unroll
        for (int r = 0; r < RM; ++r) {
          const float a = sA[ty * RM + r][c];
          acc[r][0] = fmaf(a, b.x, acc[r][0]);
          acc[r][1] = fmaf(a, b.y, acc[r][1]);
          acc[r][2] = fmaf(a, b.z, acc[r][2]);
          acc[r][3] = fmaf(a, b.w, acc[r][3]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RM; ++r) {
    const int o = row0 + ty * RM + r;
    if (o >= n_out) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int n = co_base + tx * 4 + c;
      if (n < cout) out[(long long)o * cout + n] = acc[r][c] + (bias ? bias[n] : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad: grid (kvol, S).  CTA (k, s) reduces its slice of pair list k into a Cin x Cout register
// tile (in CI_T x CO_T blocks) and adds it to gW[k] with red.global.
// ------------------------------------------------------------------------------------------------
constexpr int WP = 32;  // pairs per smem stage

template <int CI_T, int CO_T>
__global__ void __launch_bounds__(kThreads)
spconv_wgrad_kernel(const float* __restrict__ feat, const float* __restrict__ gout,
                    const int* __restrict__ pairs, const int* __restrict__ num, int pair_stride,
                    int cin, int cout, int inverse, float* __restrict__ gw) {
  ddf::pdl_sync();
  // thread tile: (CI_T*CO_T)/256 outputs, laid out TI x TJ
  constexpr int TJ = 4;
  constexpr int TXN = CO_T / TJ;           // threads along cout
  constexpr int TYN = kThreads / TXN;      // threads along cin
  constexpr int TI = CI_T / TYN;
  static_assert(TI >= 1, "tile too small");
  __shared__ float sX[WP][CI_T + 1];
  __shared__ __align__(16) float sG[WP][CO_T];
  const int k = blockIdx.x;
  const int nk = num[k];
  if (nk <= 0) return;
  const int S = gridDim.y;
  const int per = (nk + S - 1) / S;
  const int s0 = blockIdx.y * per, s1 = min(nk, s0 + per);
  if (s0 >= s1) return;
  const int* pin = pairs + ((long long)k * 2 + (inverse ? 1 : 0)) * pair_stride;
  const int* pout = pairs + ((long long)k * 2 + (inverse ? 0 : 1)) * pair_stride;
  const int tx = threadIdx.x % TXN, ty = threadIdx.x / TXN;

  for (int ci0 = 0; ci0 < cin; ci0 += CI_T) {
    for (int co0 = 0; co0 < cout; co0 += CO_T) {
      float acc[TI][TJ];
#pragma unroll
      for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[i][j] = 0.f;
      for (int p0 = s0; p0 < s1; p0 += WP) {
        __syncthreads();
        for (int e = threadIdx.x; e < WP * CI_T; e += kThreads) {
          const int p = e / CI_T, c = e % CI_T;
          float v = 0.f;
          if (p0 + p < s1 && ci0 + c < cin) v = __ldg(feat + (long long)pin[p0 + p] * cin + ci0 + c);
          sX[p][c] = v;
        }
        for (int e = threadIdx.x; e < WP * CO_T; e += kThreads) {
          const int p = e / CO_T, c = e % CO_T;
          float v = 0.f;
          if (p0 + p < s1 && co0 + c < cout) v = __ldg(gout + (long long)pout[p0 + p] * cout + co0 + c);
          sG[p][c] = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int p = 0; p < WP; ++p) {
          const float4 g = *reinterpret_cast<const float4*>(&sG[p][tx * TJ]);
#pragma unroll
          for (int i = 0; i < TI; ++i) {
            const float x = sX[p][ty * TI + i];
            acc[i][0] = fmaf(x, g.x, acc[i][0]);
            acc[i][1] = fmaf(x, g.y, acc[i][1]);
            acc[i][2] = fmaf(x, g.z, acc[i][2]);
            acc[i][3] = fmaf(x, g.w, acc[i][3]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < TI; ++i) {
        const int ci = ci0 + ty * TI + i;
        if (ci >= cin) continue;
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
          const int co = co0 + tx * TJ + j;
          if (co < cout) atomicAdd(gw + ((long long)k * cin + ci) * cout + co, acc[i][j]);
        }
      }
    }
  }
}

__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// filters [K, Cin, Cout] -> [K, Cout, Cin] (optionally rounded to tf32)
__global__ void __launch_bounds__(kThreads)
transpose_filters_kernel(const float* __restrict__ w, float* __restrict__ wt, int kvol, int cin,
                         int cout, bool round_tf32) {
  ddf::pdl_sync();
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  const long long per = (long long)cin * cout;
  if (t >= per * kvol) return;
  const int k = (int)(t / per);
  const int r = (int)(t % per);
  const int co = r / cin, ci = r % cin;
  const float v = w[(long long)k * per + (long long)ci * cout + co];
  wt[t] = round_tf32 ? tf32_rn(v) : v;
}

// dst = round-to-nearest-tf32(src): operands of the tensor-core path are made exactly
// representable, so the hardware's truncation of the low 13 mantissa bits is a no-op (unbiased
// 2^-12 relative rounding instead of a systematic 2^-11 shrink per operand)
__global__ void __launch_bounds__(kThreads)
round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n4, long long n) {
  ddf::pdl_sync();
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (t < n4) {
    float4 v = reinterpret_cast<const float4*>(src)[t];
    v.x = tf32_rn(v.x);
    v.y = tf32_rn(v.y);
    v.z = tf32_rn(v.z);
    v.w = tf32_rn(v.w);
    reinterpret_cast<float4*>(dst)[t] = v;
  }
  if (t == 0)
    for (long long i = n4 * 4; i < n; ++i) dst[i] = tf32_rn(src[i]);
}

// ---- bf16x3 operand layout -------------------------------------------------------------------------------------
// A row of C fp32 channels (C % 32 == 0) becomes C/32 blocks of 128 bytes: [32 x bf16 hi | 32 x bf16 lo] with
// hi = bf16_rn(x), lo = bf16_rn(x - hi): x = hi + lo up to 2^-17 |x|.  Same bytes per row as fp32.
__device__ __forceinline__ void split_bf16(float x, unsigned short& hi, unsigned short& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  hi = __bfloat16_as_ushort(h);
  lo = __bfloat16_as_ushort(l);
}

// src fp32 [rows, cols] -> split (block layout above) and, when rounded != nullptr, the tf32-rounded copy the
// wgrad kernels read.  One thread = 8 consecutive channels.
__global__ void __launch_bounds__(kThreads)
split_bf16x3_kernel(const float* __restrict__ src, uint8_t* __restrict__ split, float* __restrict__ rounded,
                    long long n8, int cols) {
  ddf::pdl_sync();
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (t >= n8) return;
  const long long e = t * 8;
  const long long row = e / cols;
  const int col = (int)(e % cols);
  float v[8];
  *reinterpret_cast<float4*>(v) = reinterpret_cast<const float4*>(src)[t * 2];
  *reinterpret_cast<float4*>(v + 4) = reinterpret_cast<const float4*>(src)[t * 2 + 1];
  unsigned short hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_bf16(v[i], hi[i], lo[i]);
  uint8_t* dst = split + row * cols * 4 + (col >> 5) * 128 + (col & 31) * 2;
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(dst + 64) = *reinterpret_cast<const uint4*>(lo);
  if (rounded) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = tf32_rn(v[i]);
    reinterpret_cast<float4*>(rounded)[t * 2] = *reinterpret_cast<float4*>(v);
    reinterpret_cast<float4*>(rounded)[t * 2 + 1] = *reinterpret_cast<float4*>(v + 4);
  }
}

// filters [K, Cin, Cout] -> B operand of the bf16x3 kernel: rows (k, n), contraction c in split blocks.
// transpose = true (forward): n over Cout, c over Cin; false (dgrad): n over Cin, c over Cout.
__global__ void __launch_bounds__(kThreads)
split_filters_kernel(const float* __restrict__ w, uint8_t* __restrict__ ws, int kvol, int cin, int cout,
                     bool transpose) {
  ddf::pdl_sync();
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  const long long per = (long long)cin * cout;
  if (t >= per * kvol) return;
  const int k = (int)(t / per);
  const int r = (int)(t % per);
  const int nc = transpose ? cin : cout;         // contraction length
  const int n = r / nc, c = r % nc;
  const float v = transpose ? w[(long long)k * per + (long long)c * cout + n] : w[(long long)k * per + r];
  unsigned short hi, lo;
  split_bf16(v, hi, lo);
  uint8_t* dst = ws + ((long long)k * per + (long long)n * nc) * 4 + (c >> 5) * 128 + (c & 31) * 2;
  *reinterpret_cast<unsigned short*>(dst) = hi;
  *reinterpret_cast<unsigned short*>(dst + 64) = lo;
}

// pair lists -> row-major table: table[row_of(col_sel)][k] = row_of(1-col_sel)
__global__ void __launch_bounds__(kThreads)
pairs_to_table_kernel(const int* __restrict__ pairs, const int* __restrict__ num, int pair_stride,
                      int kvol, int key_col, int* __restrict__ table) {
  ddf::pdl_sync();
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (t >= (long long)kvol * pair_stride) return;
  const int k = (int)(t / pair_stride), s = (int)(t % pair_stride);
  if (s >= num[k]) return;
  const int key = pairs[((long long)k * 2 + key_col) * pair_stride + s];
  const int val = pairs[((long long)k * 2 + (1 - key_col)) * pair_stride + s];
  table[(long long)key * kvol + k] = val;
}

// ------------------------------------------------------------------------------------------------
// Narrow layers (contraction over <= 16 channels, <= 32 output channels): the stride-1 stage of the encoders.
// At 0.075 m voxels a LiDAR sweep leaves ~1.5 rulebook pairs per output row, so the tile kernels spend
// 27 mostly-empty stages per 128 rows (0.155 ms for 240k rows, 0.41 ms for the 16->32 strided conv).
// Here a warp owns an output row: one coalesced read of its 27 table entries, a ballot of the valid
// offsets, and per valid pair Cin shuffle-broadcast FMAs against the filter bank held in shared memory
// (lane = output channel).  Full fp32, output written once, no atomics.  HBM-bound:
// 4*K (table) + 4*Cout (out) + pairs/row * 4*Cin bytes per row.
// ------------------------------------------------------------------------------------------------
template <int CI, int CO>
__global__ void __launch_bounds__(kThreads)
spconv_sparse_rows_kernel(const float* __restrict__ feat, const float* __restrict__ filt,
                          const int* __restrict__ table, const float* __restrict__ bias,
                          float* __restrict__ out, int n_out, int kvol) {
  ddf::pdl_sync();
  extern __shared__ float s_w[];   // [kvol][CI][CO]
  for (int e = threadIdx.x; e < kvol * CI * CO; e += kThreads) s_w[e] = __ldg(filt + e);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * kThreads) >> 5;
  const float b = (bias && lane < CO) ? __ldg(bias + lane) : 0.f;
  for (int o = (blockIdx.x * kThreads + threadIdx.x) >> 5; o < n_out; o += warps) {
    const int idx = lane < kvol ? __ldg(table + (long long)o * kvol + lane) : -1;
    unsigned m = __ballot_sync(0xffffffffu, idx >= 0);
    float acc = b;
    while (m) {
      const int k = __ffs(m) - 1;
      m &= m - 1;
      const int j = __shfl_sync(0xffffffffu, idx, k);
      const float f = lane < CI ? __ldg(feat + (long long)j * CI + lane) : 0.f;
      const float* wk = s_w + k * CI * CO + (lane < CO ? lane : 0);
#pragma unroll
      for (int ci = 0; ci < CI; ++ci) acc = fmaf(__shfl_sync(0xffffffffu, f, ci), wk[ci * CO], acc);
    }
    if (lane < CO) out[(long long)o * CO + lane] = acc;
  }
}

bool sparse_rows_supported(int64_t kvol, int64_t cin, int64_t cout) {
  // contraction width <= 16: at 32 input channels the tensor-core kernel is faster (measured 0.13 vs 0.25 ms)
  const bool shape = (cin == 8 || cin == 16) && (cout == 16 || cout == 32);
  return shape && kvol <= 32;
}

template <int CI, int CO>
int launch_sparse_rows(const float* feat, const float* filt, const int* table, const float* bias, float* out,
                       int64_t n_out, int kvol, cudaStream_t stream) {
  const int smem = kvol * CI * CO * 4;
  DDF_SET_SMEM_ONCE((spconv_sparse_rows_kernel<CI, CO>), 32 * CI * CO * 4);   // kvol <= 32
  long long grid = ddf::cdiv(n_out * 32, kThreads);
  const long long cap = (long long)ddf::kNumSM * (smem > 48 * 1024 ? 3 : 6);   // persistent: the filter bank is staged once per CTA
  if (grid > cap) grid = cap;
  DDF_LAUNCH_PDL((spconv_sparse_rows_kernel<CI, CO>), (unsigned)grid, kThreads, smem, stream, feat, filt, table, bias, out,
             (int)n_out, kvol);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// filt: [kvol][cin][cout] of THIS launch (dgrad passes the transposed bank)
int sparse_rows_dispatch(const float* feat, const float* filt, const int* table, const float* bias, float* out,
                         int64_t n_out, int kvol, int cin, int cout, cudaStream_t stream) {
#define DDF_SR_CASE(CI, CO) \
  if (cin == CI && cout == CO) return launch_sparse_rows<CI, CO>(feat, filt, table, bias, out, n_out, kvol, stream)
  DDF_SR_CASE(8, 16);
  DDF_SR_CASE(16, 16);
  DDF_SR_CASE(16, 32);
  DDF_SR_CASE(8, 32);
#undef DDF_SR_CASE
  ddf::set_error("sparse rows kernel: unsupported shape %d -> %d", cin, cout);
  return DDF_ERR_ARG;
}

int launch_gather_gemm(const float* feat, const float* filt, const int* table, const float* bias,
                       float* out, int64_t n_out, int kvol, int cin, int cout,
                       cudaStream_t stream) {
  if (n_out == 0) return DDF_OK;
  const unsigned grid = (unsigned)ddf::cdiv(n_out, TM);
  for (int co_base = 0; co_base < cout; co_base += 128) {
    const int rem = cout - co_base;
    if (rem > 64)
      DDF_LAUNCH(spconv_gather_gemm_kernel<128>, grid, kThreads, 0, stream, feat, filt, table, bias, out,
                                                                    (int)n_out, kvol, cin, cout, co_base);
    else if (rem > 32)
      DDF_LAUNCH(spconv_gather_gemm_kernel<64>, grid, kThreads, 0, stream, feat, filt, table, bias, out,
                                                                   (int)n_out, kvol, cin, cout, co_base);
    else if (rem > 16)
      DDF_LAUNCH(spconv_gather_gemm_kernel<32>, grid, kThreads, 0, stream, feat, filt, table, bias, out,
                                                                   (int)n_out, kvol, cin, cout, co_base);
    else
      DDF_LAUNCH(spconv_gather_gemm_kernel<16>, grid, kThreads, 0, stream, feat, filt, table, bias, out,
                                                                   (int)n_out, kvol, cin, cout, co_base);
  }
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

}  // namespace

// ---- fast path (tables from the rulebook build) ------------------------------------------------
// out [n_out, cout] = conv(features [n_in, cin], filters [K, cin, cout]) through gather_table
// [n_out, K]; bias optional ([cout] or NULL).  Fully overwrites out.
extern "C" int ddf_sparse_conv_forward(const float* features, const float* filters,
                                       const int* gather_table, const float* bias, float* out,
                                       float* filters_t_ws, int64_t n_out, int64_t n_in, int64_t kvol,
                                       int64_t cin, int64_t cout, int operand_format, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(n_out >= 0 && n_in >= 0 && kvol > 0 && cin > 0 && cout > 0, "sparse_conv_forward: bad sizes");
  if (n_out == 0) return DDF_OK;
  DDF_CHECK_ARG(features && filters && gather_table && out, "sparse_conv_forward: null pointer");
  if (operand_format == 1) {
    // features are already in the bf16 hi/lo block layout (ddf_split_bf16x3)
    DDF_CHECK_ARG(filters_t_ws && ddf::spconv_tma_supported((int)kvol, (int)cin, (int)cout),
                  "sparse_conv_forward: split operands need the multi-tile kernel (K=%lld Cin=%lld Cout=%lld)",
                  (long long)kvol, (long long)cin, (long long)cout);
    const long long nw = kvol * cin * cout;
    DDF_LAUNCH_PDL(split_filters_kernel, (unsigned)ddf::cdiv(nw, kThreads), kThreads, 0, stream, filters,
               reinterpret_cast<uint8_t*>(filters_t_ws), (int)kvol, (int)cin, (int)cout, true);
    return ddf::spconv_tma_launch(features, filters_t_ws, gather_table, bias, out, n_out, n_in, (int)kvol,
                                  (int)cin, (int)cout, false, true, stream);
  }
  if (sparse_rows_enabled() && sparse_rows_supported(kvol, cin, cout))
    return sparse_rows_dispatch(features, filters, gather_table, bias, out, n_out, (int)kvol, (int)cin, (int)cout, stream);
  if (filters_t_ws && tc_enabled() && ddf::spconv_tc_supported((int)kvol, (int)cin, (int)cout)) {
    // tensor-core path wants the filter slice K-major: Wt[k] = [cout, cin]
    const long long nw = kvol * cin * cout;
    DDF_LAUNCH_PDL(transpose_filters_kernel, (unsigned)ddf::cdiv(nw, kThreads), kThreads, 0, stream,
               filters, filters_t_ws, (int)kvol, (int)cin, (int)cout, true);
    if (tma_enabled() && ddf::spconv_tma_supported((int)kvol, (int)cin, (int)cout))
      return ddf::spconv_tma_launch(features, filters_t_ws, gather_table, bias, out, n_out, n_in, (int)kvol,
                                    (int)cin, (int)cout, gather4_enabled(), false, stream);
    return ddf::spconv_tc_launch(features, filters_t_ws, gather_table, bias, out, n_out, (int)kvol,
                                 (int)cin, (int)cout, stream);
  }
  return launch_gather_gemm(features, filters, gather_table, bias, out, n_out, (int)kvol, (int)cin,
                            (int)cout, stream);
}

// grad_in [n_in, cin] = sum_k grad_out[scatter_table[i,k]] . W[k]^T ; filters_t_ws: device scratch
// of K*cin*cout floats (receives the transposed filters).
extern "C" int ddf_sparse_conv_dgrad(const float* grad_out, const float* filters,
                                     const int* scatter_table, float* grad_in, float* filters_t_ws,
                                     int64_t n_in, int64_t n_out, int64_t kvol, int64_t cin, int64_t cout,
                                     int operand_format, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(n_in >= 0 && n_out >= -1 && kvol > 0 && cin > 0 && cout > 0, "sparse_conv_dgrad: bad sizes");
  if (n_in == 0) return DDF_OK;
  DDF_CHECK_ARG(grad_out && filters && scatter_table && grad_in && filters_t_ws,
                "sparse_conv_dgrad: null pointer");
  // dgrad contracts over Cout; filters [K, cin, cout] are already the K-major B operand [N=cin, K=cout]
  const long long nw = kvol * cin * cout;
  if (operand_format == 1) {
    DDF_CHECK_ARG(n_out >= 0 && ddf::spconv_tma_supported((int)kvol, (int)cout, (int)cin),
                  "sparse_conv_dgrad: split operands need the multi-tile kernel (K=%lld Cin=%lld Cout=%lld)",
                  (long long)kvol, (long long)cin, (long long)cout);
    DDF_LAUNCH_PDL(split_filters_kernel, (unsigned)ddf::cdiv(nw, kThreads), kThreads, 0, stream, filters,
               reinterpret_cast<uint8_t*>(filters_t_ws), (int)kvol, (int)cin, (int)cout, false);
    return ddf::spconv_tma_launch(grad_out, filters_t_ws, scatter_table, nullptr, grad_in, n_in, n_out, (int)kvol,
                                  (int)cout, (int)cin, false, true, stream);
  }
  if (sparse_rows_enabled() && sparse_rows_supported(kvol, cout, cin)) {
    DDF_LAUNCH_PDL(transpose_filters_kernel, (unsigned)ddf::cdiv(nw, kThreads), kThreads, 0, stream,
               filters, filters_t_ws, (int)kvol, (int)cin, (int)cout, false);
    return sparse_rows_dispatch(grad_out, filters_t_ws, scatter_table, nullptr, grad_in, n_in, (int)kvol, (int)cout,
                                (int)cin, stream);
  }
  if (tc_enabled() && ddf::spconv_tc_supported((int)kvol, (int)cout, (int)cin)) {
    DDF_LAUNCH_PDL(round_tf32_kernel, (unsigned)ddf::cdiv(nw / 4 + 1, kThreads), kThreads, 0, stream, filters,
               filters_t_ws, nw / 4, nw);
    if (n_out >= 0 && tma_enabled() && ddf::spconv_tma_supported((int)kvol, (int)cout, (int)cin))
      return ddf::spconv_tma_launch(grad_out, filters_t_ws, scatter_table, nullptr, grad_in, n_in, n_out,
                                    (int)kvol, (int)cout, (int)cin, gather4_enabled(), false, stream);
    return ddf::spconv_tc_launch(grad_out, filters_t_ws, scatter_table, nullptr, grad_in, n_in, (int)kvol,
                                 (int)cout, (int)cin, stream);
  }
  DDF_LAUNCH_PDL(transpose_filters_kernel, (unsigned)ddf::cdiv(nw, kThreads), kThreads, 0, stream,
             filters, filters_t_ws, (int)kvol, (int)cin, (int)cout, false);
  // dgrad is a conv with Cin<->Cout swapped through the transposed table
  return launch_gather_gemm(grad_out, filters_t_ws, scatter_table, nullptr, grad_in, n_in,
                            (int)kvol, (int)cout, (int)cin, stream);
}

// grad_filters [K, cin, cout] (zeroed inside) from the reference-format pair lists.
static int sparse_conv_wgrad_impl(const float* features, const float* grad_out,
                                  const int* indice_pairs, const int* indice_num,
                                  int64_t pair_stride, float* grad_filters, int64_t kvol,
                                  int64_t cin, int64_t cout, int inverse, bool allow_tc, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(kvol > 0 && cin > 0 && cout > 0 && pair_stride >= 0, "sparse_conv_wgrad: bad sizes");
  DDF_CHECK_ARG(grad_filters != nullptr, "sparse_conv_wgrad: null grad_filters");
  DDF_CUDA(cudaMemsetAsync(grad_filters, 0, sizeof(float) * (size_t)(kvol * cin * cout), stream));
  if (pair_stride == 0) return DDF_OK;
  DDF_CHECK_ARG(features && grad_out && indice_pairs && indice_num, "sparse_conv_wgrad: null pointer");
  if (allow_tc && tc_enabled() && ddf::spconv_wgrad_tc_supported((int)kvol, (int)cin, (int)cout))
    return ddf::spconv_wgrad_tc_launch(features, grad_out, indice_pairs, indice_num, pair_stride,
                                       grad_filters, (int)kvol, (int)cin, (int)cout, inverse, stream);
  // enough (k, slice) CTAs for ~4 waves; slices bounded so each still has >= ~256 pairs
  int S = (int)ddf::cdiv(4 * ddf::kNumSM, kvol);
  const int maxS = (int)ddf::cdiv(pair_stride, 256);
  if (S > maxS) S = maxS;
  if (S < 1) S = 1;
  dim3 grid((unsigned)kvol, (unsigned)S);
  if (cin <= 32 && cout <= 32)
    DDF_LAUNCH((spconv_wgrad_kernel<32, 32>), grid, kThreads, 0, stream, 
        features, grad_out, indice_pairs, indice_num, (int)pair_stride, (int)cin, (int)cout, inverse, grad_filters);
  else
    DDF_LAUNCH((spconv_wgrad_kernel<64, 64>), grid, kThreads, 0, stream, 
        features, grad_out, indice_pairs, indice_num, (int)pair_stride, (int)cin, (int)cout, inverse, grad_filters);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

extern "C" int ddf_sparse_conv_wgrad(const float* features, const float* grad_out,
                                     const int* indice_pairs, const int* indice_num,
                                     int64_t pair_stride, float* grad_filters, int64_t kvol,
                                     int64_t cin, int64_t cout, int inverse, void* stream_) {
  return sparse_conv_wgrad_impl(features, grad_out, indice_pairs, indice_num, pair_stride, grad_filters, kvol, cin,
                                cout, inverse, true, stream_);
}

// Runtime switch of the tensor-core conv kernels; returns the previous setting.
extern "C" int ddf_set_tensor_cores(int on) {
  const int prev = tc_mode_state();
  g_tc_state = on < 0 ? 0 : (on > 4 ? 4 : on);
  return prev;
}

// grad_filters [K, cin, cout] (zeroed inside) through the gather table [n_out, K] of the forward conv:
// the output rows are walked once, gout rows are read densely, only feature rows are gathered.
// Returns DDF_ERR_ARG when the layer shape is not supported (callers fall back to the pair lists).
extern "C" int ddf_sparse_conv_wgrad_table(const float* features, const float* grad_out,
                                           const int* gather_table, float* grad_filters, int64_t n_out,
                                           int64_t n_in, int64_t kvol, int64_t cin, int64_t cout,
                                           void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(n_out >= 0 && n_in >= 0 && kvol > 0 && cin > 0 && cout > 0, "sparse_conv_wgrad_table: bad sizes");
  DDF_CHECK_ARG(grad_filters != nullptr, "sparse_conv_wgrad_table: null grad_filters");
  DDF_CHECK_ARG(tc_enabled() && ddf::spconv_wgrad_table_supported((int)kvol, (int)cin, (int)cout),
                "sparse_conv_wgrad_table: unsupported layer shape K=%lld Cin=%lld Cout=%lld", (long long)kvol,
                (long long)cin, (long long)cout);
  DDF_CUDA(cudaMemsetAsync(grad_filters, 0, sizeof(float) * (size_t)(kvol * cin * cout), stream));
  if (n_out == 0) return DDF_OK;
  DDF_CHECK_ARG(features && grad_out && gather_table, "sparse_conv_wgrad_table: null pointer");
  return ddf::spconv_wgrad_table_launch(features, grad_out, gather_table, grad_filters, n_out, n_in, (int)kvol,
                                        (int)cin, (int)cout, stream);
}

// Which of the three conv kernels of a (kvol, cin, cout) layer run on the tensor cores:
// bit 0 forward, bit 1 dgrad, bit 2 wgrad (pair lists), bit 3 wgrad (table); bit 4 / bit 5: forward / dgrad take
// their operands in the bf16 hi/lo block layout (operand_format 1).  Callers use it to prepare operands.
extern "C" int ddf_sparse_conv_tc_mode(int64_t kvol, int64_t cin, int64_t cout) {
  if (!tc_enabled()) return 0;
  int m = 0;
  const bool rows_f = sparse_rows_enabled() && sparse_rows_supported(kvol, cin, cout);
  const bool rows_d = sparse_rows_enabled() && sparse_rows_supported(kvol, cout, cin);
  if (!rows_f && ddf::spconv_tc_supported((int)kvol, (int)cin, (int)cout)) m |= 1;
  if (!rows_d && ddf::spconv_tc_supported((int)kvol, (int)cout, (int)cin)) m |= 2;
  if (ddf::spconv_wgrad_tc_supported((int)kvol, (int)cin, (int)cout)) m |= 4;
  if (tc_mode_state() != 2 && ddf::spconv_wgrad_table_supported((int)kvol, (int)cin, (int)cout)) m |= 8;
  if (bf16x3_enabled()) {
    if ((m & 1) && ddf::spconv_tma_supported((int)kvol, (int)cin, (int)cout)) m |= 16;
    if ((m & 2) && ddf::spconv_tma_supported((int)kvol, (int)cout, (int)cin)) m |= 32;
  }
  return m;
}

// dst[i] = src[i] rounded to the nearest tf32 (dst may alias src); 16-byte aligned pointers.
extern "C" int ddf_round_tf32(const float* src, float* dst, int64_t n, void* stream_) {
  DDF_CHECK_ARG(n >= 0, "round_tf32: bad size");
  if (n == 0) return DDF_OK;
  DDF_CHECK_ARG(src && dst, "round_tf32: null pointer");
  DDF_LAUNCH_PDL(round_tf32_kernel, (unsigned)ddf::cdiv(n / 4 + 1, kThreads), kThreads, 0,
             (cudaStream_t)stream_, src, dst, n / 4, n);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// src fp32 [rows, cols] (cols % 32 == 0) -> split: the bf16 hi/lo block layout of the bf16x3 kernels (rows*cols*4
// bytes); rounded (optional, may be NULL): the tf32-rounded copy for the wgrad kernels.
extern "C" int ddf_split_bf16x3(const float* src, void* split, float* rounded, int64_t rows, int64_t cols,
                                void* stream_) {
  DDF_CHECK_ARG(rows >= 0 && cols > 0 && cols % 32 == 0, "split_bf16x3: cols must be a multiple of 32");
  if (rows == 0) return DDF_OK;
  DDF_CHECK_ARG(src && split, "split_bf16x3: null pointer");
  const long long n8 = rows * cols / 8;
  DDF_LAUNCH_PDL(split_bf16x3_kernel, (unsigned)ddf::cdiv(n8, kThreads), kThreads, 0, (cudaStream_t)stream_, src,
             reinterpret_cast<uint8_t*>(split), rounded, n8, (int)cols);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// ---- drop-in path (reference-format rulebook only) ---------------------------------------------
// Same contract as sparse_conv_ext.indice_conv_fp32 (spconv_ops.h:260-361): the gather table is
// rebuilt from the pair lists into table_ws ([n_out, K] int32 scratch).
extern "C" int ddf_indice_conv(const float* features, const float* filters, const int* indice_pairs,
                               const int* indice_num, int64_t pair_stride, float* out,
                               int64_t n_out, int64_t kvol, int64_t cin, int64_t cout, int inverse,
                               int subm, int* table_ws, void* stream_) {
  (void)subm;
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(n_out >= 0 && kvol > 0 && cin > 0 && cout > 0, "indice_conv: bad sizes");
  if (n_out == 0) return DDF_OK;
  DDF_CHECK_ARG(features && filters && indice_pairs && indice_num && out && table_ws,
                "indice_conv: null pointer");
  DDF_CUDA(cudaMemsetAsync(table_ws, 0xff, sizeof(int) * (size_t)(n_out * kvol), stream));
  const long long np = kvol * pair_stride;
  if (np > 0)
    DDF_LAUNCH(pairs_to_table_kernel, (unsigned)ddf::cdiv(np, kThreads), kThreads, 0, stream, 
        indice_pairs, indice_num, (int)pair_stride, (int)kvol, inverse ? 0 : 1, table_ws);
  return launch_gather_gemm(features, filters, table_ws, nullptr, out, n_out, (int)kvol, (int)cin,
                            (int)cout, stream);
}

// Same contract as sparse_conv_ext.indice_conv_backward_fp32 (spconv_ops.h:363-456).
// table_ws: [n_in, K] int32 scratch; filters_t_ws: K*cin*cout floats scratch.
extern "C" int ddf_indice_conv_backward(const float* features, const float* filters,
                                        const float* grad_out, const int* indice_pairs,
                                        const int* indice_num, int64_t pair_stride, float* grad_in,
                                        float* grad_filters, int64_t n_in, int64_t kvol,
                                        int64_t cin, int64_t cout, int inverse, int subm,
                                        int* table_ws, float* filters_t_ws, void* stream_) {
  (void)subm;
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(n_in >= 0 && kvol > 0 && cin > 0 && cout > 0, "indice_conv_backward: bad sizes");
  // the reference-ABI entry points are the indice_conv*_fp32 contract: fp32 SIMT kernels in forward AND
  // backward (the tensor-core kernels, which need prepared operands, are reached through ddf_sparse_conv_*)
  int rc = sparse_conv_wgrad_impl(features, grad_out, indice_pairs, indice_num, pair_stride,
                                  grad_filters, kvol, cin, cout, inverse, false, stream_);
  if (rc || n_in == 0) return rc;
  DDF_CHECK_ARG(grad_in && table_ws && filters_t_ws, "indice_conv_backward: null pointer");
  DDF_CUDA(cudaMemsetAsync(table_ws, 0xff, sizeof(int) * (size_t)(n_in * kvol), stream));
  const long long np = kvol * pair_stride;
  if (np > 0)
    DDF_LAUNCH(pairs_to_table_kernel, (unsigned)ddf::cdiv(np, kThreads), kThreads, 0, stream, 
        indice_pairs, indice_num, (int)pair_stride, (int)kvol, inverse ? 1 : 0, table_ws);
  const long long nw = kvol * cin * cout;
  DDF_LAUNCH_PDL(transpose_filters_kernel, (unsigned)ddf::cdiv(nw, kThreads), kThreads, 0, stream,
             filters, filters_t_ws, (int)kvol, (int)cin, (int)cout, false);
  return launch_gather_gemm(grad_out, filters_t_ws, table_ws, nullptr, grad_in, n_in, (int)kvol, (int)cout,
                            (int)cin, stream);
}

// ---- dense(): sparse [N, C] + indices [N, 4] -> dense NCDHW ------------------------------------
// Replaces SparseConvTensor.dense() = scatter_nd + permute(...).contiguous()
//   <TF>/ops/spconv/structure.py:5-18, 55-64 (zeros + index_put + a full permute copy).
namespace {
__global__ void __launch_bounds__(kThreads)
dense_scatter_kernel(const float* __restrict__ feat, const int* __restrict__ indices, int n, int C,
                     int D, int H, int W, float* __restrict__ out) {
  ddf::pdl_sync();
  // one warp per voxel row keeps the feature read coalesced; the NCDHW write is a C-strided scatter
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  const int row = (int)(t >> 5), lane = (int)(t & 31);
  if (row >= n) return;
  const int4 c = reinterpret_cast<const int4*>(indices)[row];
  const long long plane = (long long)D * H * W;
  const long long base = (long long)c.x * C * plane + ((long long)c.y * H + c.z) * W + c.w;
  for (int ch = lane; ch < C; ch += 32) out[base + ch * plane] = feat[(long long)row * C + ch];
}

__global__ void __launch_bounds__(kThreads)
dense_gather_kernel(const float* __restrict__ gdense, const int* __restrict__ indices, int n, int C,
                    int D, int H, int W, float* __restrict__ gfeat) {
  ddf::pdl_sync();
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  const int row = (int)(t >> 5), lane = (int)(t & 31);
  if (row >= n) return;
  const int4 c = reinterpret_cast<const int4*>(indices)[row];
  const long long plane = (long long)D * H * W;
  const long long base = (long long)c.x * C * plane + ((long long)c.y * H + c.z) * W + c.w;
  for (int ch = lane; ch < C; ch += 32) gfeat[(long long)row * C + ch] = gdense[base + ch * plane];
}
}  // namespace

// out [B, C, D, H, W] (zeroed inside, then scattered).
extern "C" int ddf_sparse_to_dense(const float* features, const int* indices, float* out,
                                   int64_t n, int64_t C, int64_t B, int64_t D, int64_t H,
                                   int64_t W, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(n >= 0 && C > 0 && B > 0 && D > 0 && H > 0 && W > 0, "sparse_to_dense: bad sizes");
  DDF_CHECK_ARG(out != nullptr, "sparse_to_dense: null out");
  DDF_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)(B * C * D * H * W), stream));
  if (n == 0) return DDF_OK;
  DDF_CHECK_ARG(features && indices, "sparse_to_dense: null pointer");
  DDF_LAUNCH_PDL(dense_scatter_kernel, (unsigned)ddf::cdiv(n * 32, kThreads), kThreads, 0, stream, 
      features, indices, (int)n, (int)C, (int)D, (int)H, (int)W, out);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// backward of dense(): grad_features [n, C] = grad_dense at the active cells.
extern "C" int ddf_dense_to_sparse(const float* grad_dense, const int* indices,
                                   float* grad_features, int64_t n, int64_t C, int64_t B,
                                   int64_t D, int64_t H, int64_t W, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(n >= 0 && C > 0 && B > 0 && D > 0 && H > 0 && W > 0, "dense_to_sparse: bad sizes");
  if (n == 0) return DDF_OK;
  DDF_CHECK_ARG(grad_dense && indices && grad_features, "dense_to_sparse: null pointer");
  DDF_LAUNCH_PDL(dense_gather_kernel, (unsigned)ddf::cdiv(n * 32, kThreads), kThreads, 0, stream, 
      grad_dense, indices, (int)n, (int)C, (int)D, (int)H, (int)W, grad_features);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// ---- BEV hand-off: sparse [N, C] + indices [N, 4] -> the 2-D backbone's input in bf16, channels-last ----------
// The reference hands dense() (fp32 NCDHW) viewed as [B, C*D, H, W] to SECOND (sparse_encoder.py:366-367,
// backbones/second.py).  For a bf16 channels-last 2-D backbone the same map is written directly as the physical
// layout [B, H, W, C*D] (channel index c * D + z), half the bytes and no permute.
namespace {
__global__ void __launch_bounds__(kThreads)
bev_nhwc_bf16_scatter_kernel(const float* __restrict__ feat, const int* __restrict__ indices, int n, int C, int D,
                             int H, int W, __nv_bfloat16* __restrict__ out) {
  ddf::pdl_sync();
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  const int row = (int)(t >> 5), lane = (int)(t & 31);
  if (row >= n) return;
  const int4 c = reinterpret_cast<const int4*>(indices)[row];      // b, z, y, x
  __nv_bfloat16* base = out + (((long long)c.x * H + c.z) * W + c.w) * ((long long)C * D) + c.y;
  for (int ch = lane; ch < C; ch += 32) base[(long long)ch * D] = __float2bfloat16_rn(feat[(long long)row * C + ch]);
}

__global__ void __launch_bounds__(kThreads)
bev_nhwc_bf16_gather_kernel(const __nv_bfloat16* __restrict__ gdense, const int* __restrict__ indices, int n, int C,
                            int D, int H, int W, float* __restrict__ gfeat) {
  ddf::pdl_sync();
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  const int row = (int)(t >> 5), lane = (int)(t & 31);
  if (row >= n) return;
  const int4 c = reinterpret_cast<const int4*>(indices)[row];
  const __nv_bfloat16* base = gdense + (((long long)c.x * H + c.z) * W + c.w) * ((long long)C * D) + c.y;
  for (int ch = lane; ch < C; ch += 32) gfeat[(long long)row * C + ch] = __bfloat162float(base[(long long)ch * D]);
}
}  // namespace

// out: bf16 [B, H, W, C*D] (zeroed inside) = channels-last storage of the logical [B, C*D, H, W] BEV map
extern "C" int ddf_sparse_to_bev_nhwc_bf16(const float* features, const int* indices, void* out, int64_t n,
                                           int64_t C, int64_t B, int64_t D, int64_t H, int64_t W, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(n >= 0 && C > 0 && B > 0 && D > 0 && H > 0 && W > 0, "sparse_to_bev_nhwc_bf16: bad sizes");
  DDF_CHECK_ARG(out != nullptr, "sparse_to_bev_nhwc_bf16: null out");
  DDF_CUDA(cudaMemsetAsync(out, 0, 2 * (size_t)(B * C * D * H * W), stream));
  if (n == 0) return DDF_OK;
  DDF_CHECK_ARG(features && indices, "sparse_to_bev_nhwc_bf16: null pointer");
  DDF_LAUNCH_PDL(bev_nhwc_bf16_scatter_kernel, (unsigned)ddf::cdiv(n * 32, kThreads), kThreads, 0, stream, features,
             indices, (int)n, (int)C, (int)D, (int)H, (int)W, reinterpret_cast<__nv_bfloat16*>(out));
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

extern "C" int ddf_bev_nhwc_bf16_to_sparse(const void* grad_dense, const int* indices, float* grad_features,
                                           int64_t n, int64_t C, int64_t B, int64_t D, int64_t H, int64_t W,
                                           void* stream_) {
  DDF_CHECK_ARG(n >= 0 && C > 0 && B > 0 && D > 0 && H > 0 && W > 0, "bev_nhwc_bf16_to_sparse: bad sizes");
  if (n == 0) return DDF_OK;
  DDF_CHECK_ARG(grad_dense && indices && grad_features, "bev_nhwc_bf16_to_sparse: null pointer");
  DDF_LAUNCH_PDL(bev_nhwc_bf16_gather_kernel, (unsigned)ddf::cdiv(n * 32, kThreads), kThreads, 0, (cudaStream_t)stream_,
             reinterpret_cast<const __nv_bfloat16*>(grad_dense), indices, (int)n, (int)C, (int)D, (int)H, (int)W,
             grad_features);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}
