// Camera-side input projection of the fusion encoder in the token-major ("rows") layout.
//
// The reference runs input_proj = Conv2d(k=1) + GroupNorm(32, d_model) on NCHW camera maps and then flattens /
// transposes the result to [B', H*W, d_model] for the encoder (<proj>/models/model_utils/actr.py:131-187,
// actr_transformer.py:255-264); the per-query camera feature goes through Conv1d(k=1) + GroupNorm between two more
// transposes (actr.py:150-158).  On B200 those layout changes and the NCHW GroupNorm were 1.5 ms of a 34 ms step: two
// full-map transposing copies forward and backward, a 256-channel gather with a 90 KB stride between channels, and
// ATen's GroupNorm at a third of the HBM rate.  Here the maps are turned into rows ONCE (ddf_nchw_to_rows, from fp32
// or bf16), the 1x1 convolution is a row-major GEMM, and GroupNorm runs on rows:
//   forward : gn_rows_stats_kernel (per (sample, group) sum / sum of squares, fp64 atomics) -> gn_rows_finalize_kernel
//             -> gn_rows_apply_kernel
//   backward: gn_rows_bwd_reduce_kernel (per (sample, group) sums of dy*w and dy*w*xhat, per-channel grad_weight /
//             grad_bias) -> gn_rows_bwd_apply_kernel
// HBM-bound streaming kernels: a thread owns one 16-byte vector of a row = the 4 channels of ONE group (C / G == 4,
// the shape of every shipped config: GroupNorm(32, 128)).  Algorithmic bytes: forward 4*N*L*C*(2 reads + 1 write),
// backward 4*N*L*C*(2 + 2 reads + 1 write).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {
constexpr int kThreads = 256;

// ---- [N, C, HW] (fp32 or bf16) -> [N, HW, C] fp32: 32 x 32 tiles through shared memory ------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_rows_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, int HW) {
  ddf::pdl_sync();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const T* s = src + (long long)n * C * HW;
  float* d = dst + (long long)n * HW * C;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int c = c0 + ty + j, p = p0 + tx;
    tile[ty + j][tx] = (c < C && p < HW) ? to_f32<T>(s[(long long)c * HW + p]) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int p = p0 + ty + j, c = c0 + tx;
    if (p < HW && c < C) d[(long long)p * C + c] = tile[tx][ty + j];
  }
}

// ---- GroupNorm on rows ---------------------------------------------------------------------------------------
// x [N, L, C], G = C / 4 groups; ws [N * G * 2] doubles (zeroed by the launcher).
__global__ void __launch_bounds__(kThreads)
gn_rows_stats_kernel(const float* __restrict__ x, double* __restrict__ ws, int L, int C) {
  ddf::pdl_sync();
  const int tpr = C >> 2, rpb = kThreads / tpr;
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  const int n = blockIdx.y;
  const float* px = x + (long long)n * L * C + cg * 4;
  float s = 0.f, q = 0.f;
  const int step = gridDim.x * rpb;
  int r = blockIdx.x * rpb + rl;
  if (rl < rpb) {
    for (; r + 3 * step < L; r += 4 * step) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ldg4(px + (long long)(r + u * step) * C);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
        q = fmaf(v[u].x, v[u].x, fmaf(v[u].y, v[u].y, fmaf(v[u].z, v[u].z, fmaf(v[u].w, v[u].w, q))));
      }
    }
    for (; r < L; r += step) {
      const float4 v = ldg4(px + (long long)r * C);
      s += (v.x + v.y) + (v.z + v.w);
      q = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, q))));
    }
  }
  __shared__ double sm[2][kThreads];
  sm[0][threadIdx.x] = (double)s;
  sm[1][threadIdx.x] = (double)q;
  __syncthreads();
  if (rl == 0) {
    double a = sm[0][threadIdx.x], b = sm[1][threadIdx.x];
    for (int j = 1; j < rpb; ++j) {
      a += sm[0][j * tpr + cg];
      b += sm[1][j * tpr + cg];
    }
    double* w = ws + ((long long)n * tpr + cg) * 2;
    atomicAdd(w, a);
    atomicAdd(w + 1, b);
  }
}

__global__ void __launch_bounds__(kThreads)
gn_rows_finalize_kernel(const double* __restrict__ ws, float* __restrict__ mean, float* __restrict__ rstd, int NG,
                        double count, float eps) {
  ddf::pdl_sync();
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= NG) return;
  const double m = ws[2 * i] / count;
  double var = ws[2 * i + 1] / count - m * m;   // biased, as torch.nn.GroupNorm
  if (var < 0.0) var = 0.0;
  mean[i] = (float)m;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void __launch_bounds__(kThreads)
gn_rows_apply_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                     const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ y,
                     long long n4, int L, int C) {
  ddf::pdl_sync();
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n4) return;
  const int tpr = C >> 2;
  const int cg = (int)(i % tpr);
  const long long n = i / ((long long)L * tpr);
  const float m = __ldg(mean + n * tpr + cg), rs = __ldg(rstd + n * tpr + cg);
  const float4 v = ldg4(x + i * 4);
  const float4 ww = w ? ldg4(w + cg * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 bb = b ? ldg4(b + cg * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 o;
  o.x = fmaf((v.x - m) * rs, ww.x, bb.x);
  o.y = fmaf((v.y - m) * rs, ww.y, bb.y);
  o.z = fmaf((v.z - m) * rs, ww.z, bb.z);
  o.w = fmaf((v.w - m) * rs, ww.w, bb.w);
  *reinterpret_cast<float4*>(y + i * 4) = o;
}

// per (sample, group): A = sum dy*w, B = sum dy*w*xhat (ws, fp64 atomics); per channel: gw += dy*xhat, gb += dy
__global__ void __launch_bounds__(kThreads)
gn_rows_bwd_reduce_kernel(const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ w,
                          const float* __restrict__ mean, const float* __restrict__ rstd, double* __restrict__ ws,
                          float* __restrict__ gw, float* __restrict__ gb, int L, int C) {
  ddf::pdl_sync();
  const int tpr = C >> 2, rpb = kThreads / tpr;
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  const int n = blockIdx.y;
  const long long base = (long long)n * L * C + cg * 4;
  const float m = __ldg(mean + n * tpr + cg), rs = __ldg(rstd + n * tpr + cg);
  const float4 ww = w ? ldg4(w + cg * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
  float a = 0.f, bsum = 0.f;
  float4 dw = make_float4(0.f, 0.f, 0.f, 0.f), db = dw;
  auto add = [&](const float4 g, const float4 v) {
    const float4 xh = make_float4((v.x - m) * rs, (v.y - m) * rs, (v.z - m) * rs, (v.w - m) * rs);
    const float4 gwv = make_float4(g.x * ww.x, g.y * ww.y, g.z * ww.z, g.w * ww.w);
    a += (gwv.x + gwv.y) + (gwv.z + gwv.w);
    bsum = fmaf(gwv.x, xh.x, fmaf(gwv.y, xh.y, fmaf(gwv.z, xh.z, fmaf(gwv.w, xh.w, bsum))));
    dw.x = fmaf(g.x, xh.x, dw.x); dw.y = fmaf(g.y, xh.y, dw.y); dw.z = fmaf(g.z, xh.z, dw.z); dw.w = fmaf(g.w, xh.w, dw.w);
    db.x += g.x; db.y += g.y; db.z += g.z; db.w += g.w;
  };
  const int step = gridDim.x * rpb;
  int r = blockIdx.x * rpb + rl;
  if (rl < rpb) {
    for (; r + 3 * step < L; r += 4 * step) {
      float4 g[4], v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long o = base + (long long)(r + u * step) * C;
        g[u] = ldg4(gy + o);
        v[u] = ldg4(x + o);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) add(g[u], v[u]);
    }
    for (; r < L; r += step) {
      const long long o = base + (long long)r * C;
      add(ldg4(gy + o), ldg4(x + o));
    }
  }
  __shared__ float sm[10][kThreads];
  sm[0][threadIdx.x] = a; sm[1][threadIdx.x] = bsum;
  sm[2][threadIdx.x] = dw.x; sm[3][threadIdx.x] = dw.y; sm[4][threadIdx.x] = dw.z; sm[5][threadIdx.x] = dw.w;
  sm[6][threadIdx.x] = db.x; sm[7][threadIdx.x] = db.y; sm[8][threadIdx.x] = db.z; sm[9][threadIdx.x] = db.w;
  __syncthreads();
  if (rl == 0) {
    float t[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) t[k] = sm[k][threadIdx.x];
    for (int j = 1; j < rpb; ++j) {
#pragma unroll
      for (int k = 0; k < 10; ++k) t[k] += sm[k][j * tpr + cg];
    }
    double* wsn = ws + ((long long)n * tpr + cg) * 2;
    atomicAdd(wsn, (double)t[0]);
    atomicAdd(wsn + 1, (double)t[1]);
    if (gw) red_add_v4(gw + cg * 4, t[2], t[3], t[4], t[5]);
    if (gb) red_add_v4(gb + cg * 4, t[6], t[7], t[8], t[9]);
  }
}

// gx = rstd * (dy*w - A / cnt - xhat * B / cnt)
__global__ void __launch_bounds__(kThreads)
gn_rows_bwd_apply_kernel(const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ w,
                         const float* __restrict__ mean, const float* __restrict__ rstd,
                         const double* __restrict__ ws, float* __restrict__ gx, long long n4, int L, int C,
                         float inv_count) {
  ddf::pdl_sync();
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n4) return;
  const int tpr = C >> 2;
  const int cg = (int)(i % tpr);
  const long long n = i / ((long long)L * tpr);
  const float m = __ldg(mean + n * tpr + cg), rs = __ldg(rstd + n * tpr + cg);
  const float c1 = (float)ws[(n * tpr + cg) * 2] * inv_count, c2 = (float)ws[(n * tpr + cg) * 2 + 1] * inv_count;
  const float4 g = ldg4(gy + i * 4), v = ldg4(x + i * 4);
  const float4 ww = w ? ldg4(w + cg * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
  float4 o;
  o.x = rs * (g.x * ww.x - c1 - (v.x - m) * rs * c2);
  o.y = rs * (g.y * ww.y - c1 - (v.y - m) * rs * c2);
  o.z = rs * (g.z * ww.z - c1 - (v.z - m) * rs * c2);
  o.w = rs * (g.w * ww.w - c1 - (v.w - m) * rs * c2);
  *reinterpret_cast<float4*>(gx + i * 4) = o;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline unsigned chunks_for(int64_t N, int64_t L, int rpb) {
  long long c = ddf::cdiv(4 * ddf::kNumSM, N);                 // about 4 CTAs per SM over all samples
  const long long most = ddf::cdiv(L, (long long)rpb * 4);     // at least 4 rows per thread
  if (c > most) c = most;
  return (unsigned)(c < 1 ? 1 : c);
}
}  // namespace

extern "C" int ddf_group_norm_rows_supported(int64_t C, int64_t G) {
  return G > 0 && C == 4 * G && C / 4 <= kThreads && kThreads % (C / 4) == 0;
}

// src [N, C, HW] (dtype 0: fp32, 1: bf16) -> dst [N, HW, C] fp32
extern "C" int ddf_nchw_to_rows(const void* src, int dtype, float* dst, int64_t N, int64_t C, int64_t HW, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(N >= 0 && C > 0 && HW > 0 && N < 65536 && C <= 65535ll * 32 && HW < (1ll << 31) && (dtype == 0 || dtype == 1),
                "nchw_to_rows: bad arguments (N < 65536, C <= 2097120)");
  if (N == 0) return DDF_OK;
  DDF_CHECK_ARG(src && dst, "nchw_to_rows: null pointer");
  const dim3 grid((unsigned)ddf::cdiv(HW, 32), (unsigned)ddf::cdiv(C, 32), (unsigned)N);
  if (dtype == 0)
    DDF_LAUNCH_PDL(nchw_to_rows_kernel<float>, grid, 256, 0, stream, (const float*)src, dst, (int)C, (int)HW);
  else
    DDF_LAUNCH_PDL(nchw_to_rows_kernel<__nv_bfloat16>, grid, 256, 0, stream, (const __nv_bfloat16*)src, dst, (int)C,
                   (int)HW);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// y [N, L, C] = GroupNorm(G = C / 4 groups) of x over (L, C / G) per sample, affine w / b [C] (may be NULL).
// mean / rstd [N, G] are saved for backward; ws: N * G * 2 doubles of scratch.
extern "C" int ddf_group_norm_rows_forward(const float* x, const float* w, const float* b, float* y, float* mean,
                                           float* rstd, void* ws, int64_t N, int64_t L, int64_t C, int64_t G,
                                           float eps, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(N >= 0 && L >= 0 && N < 65536 && ddf_group_norm_rows_supported(C, G),
                "group_norm_rows: needs C == 4 * G and C / 4 dividing %d (C=%lld G=%lld)", kThreads, (long long)C,
                (long long)G);
  if (N == 0 || L == 0) return DDF_OK;
  DDF_CHECK_ARG(x && y && mean && rstd && ws, "group_norm_rows: null pointer");
  DDF_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(w) && aligned16(b) && aligned16(ws),
                "group_norm_rows: misaligned pointer");
  const int tpr = (int)(C / 4), rpb = kThreads / tpr;
  DDF_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * N * G, stream));
  const dim3 grid(chunks_for(N, L, rpb), (unsigned)N);
  DDF_LAUNCH_PDL(gn_rows_stats_kernel, grid, kThreads, 0, stream, x, (double*)ws, (int)L, (int)C);
  DDF_LAUNCH_PDL(gn_rows_finalize_kernel, (unsigned)ddf::cdiv(N * G, kThreads), kThreads, 0, stream, (const double*)ws, mean,
                 rstd, (int)(N * G), (double)L * 4.0, eps);
  const long long n4 = N * L * tpr;
  DDF_LAUNCH_PDL(gn_rows_apply_kernel, (unsigned)ddf::cdiv(n4, kThreads), kThreads, 0, stream, x, w, b, (const float*)mean,
                 (const float*)rstd, y, n4, (int)L, (int)C);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// grad_x [N, L, C]; grad_w / grad_b [C] are overwritten (may be NULL).
extern "C" int ddf_group_norm_rows_backward(const float* grad_y, const float* x, const float* w, const float* mean,
                                            const float* rstd, float* grad_x, float* grad_w, float* grad_b, void* ws,
                                            int64_t N, int64_t L, int64_t C, int64_t G, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(N >= 0 && L >= 0 && N < 65536 && ddf_group_norm_rows_supported(C, G),
                "group_norm_rows_backward: needs C == 4 * G and C / 4 dividing %d (C=%lld G=%lld)", kThreads,
                (long long)C, (long long)G);
  if (grad_w) DDF_CUDA(cudaMemsetAsync(grad_w, 0, sizeof(float) * C, stream));
  if (grad_b) DDF_CUDA(cudaMemsetAsync(grad_b, 0, sizeof(float) * C, stream));
  if (N == 0 || L == 0) return DDF_OK;
  DDF_CHECK_ARG(grad_y && x && mean && rstd && grad_x && ws, "group_norm_rows_backward: null pointer");
  DDF_CHECK_ARG(aligned16(grad_y) && aligned16(x) && aligned16(w) && aligned16(grad_x) && aligned16(grad_w) &&
                    aligned16(grad_b) && aligned16(ws),
                "group_norm_rows_backward: misaligned pointer");
  const int tpr = (int)(C / 4), rpb = kThreads / tpr;
  DDF_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * N * G, stream));
  const dim3 grid(chunks_for(N, L, rpb), (unsigned)N);
  DDF_LAUNCH_PDL(gn_rows_bwd_reduce_kernel, grid, kThreads, 0, stream, grad_y, x, w, mean, rstd, (double*)ws, grad_w, grad_b,
                 (int)L, (int)C);
  const long long n4 = N * L * tpr;
  DDF_LAUNCH_PDL(gn_rows_bwd_apply_kernel, (unsigned)ddf::cdiv(n4, kThreads), kThreads, 0, stream, grad_y, x, w, mean, rstd,
                 (const double*)ws, grad_x, n4, (int)L, (int)C, (float)(1.0 / ((double)L * 4.0)));
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}
