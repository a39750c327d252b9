// Fused elementwise kernels of the 3D-DF encoder layers (fp32, sm_100a), replacing chains of separate
// PyTorch kernels in <proj>/models/model_utils/actr_transformer.py:
//
//   forward_ffn (:383-397):  linear2(dropout(relu(linear1(x))))
//       -> ddf_bias_relu_dropout_forward / _backward on the [tokens, d_ffn] hidden activation (598 MB per
//          FFN at the TransFusion config): bias add + ReLU + dropout in ONE in-place pass instead of three
//          read+write passes; backward needs no mask tensor: out != 0  <=>  kept and positive.
//   norm(src + dropout(src2)) (:385, :391, :406):
//       -> ddf_add_dropout_layer_norm_forward / _backward: residual add + dropout + LayerNorm, a warp per
//          token row (torch's LayerNorm kernel takes 224 us for 146k x 128 rows; this is one 4-pass stream).
//
// Dropout uses a counter-based hash of (seed, element index): nothing is stored, backward recomputes
// the keep decision.  All kernels are HBM-bound streaming kernels with 16-byte accesses.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// 64-bit mix (splitmix64 finaliser) of (seed, index of a group of 4 elements) -> four 16-bit uniforms
__device__ __forceinline__ unsigned long long mix(unsigned long long seed, unsigned long long i) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// keep flags of the 4 elements of vector i (bit j = element j kept); thr = p * 65536
__device__ __forceinline__ unsigned keep4(unsigned long long seed, unsigned long long i, unsigned thr) {
  const unsigned long long r = mix(seed, i);
  return ((unsigned)(r & 0xffff) >= thr ? 1u : 0u) | ((unsigned)((r >> 16) & 0xffff) >= thr ? 2u : 0u) |
         ((unsigned)((r >> 32) & 0xffff) >= thr ? 4u : 0u) | ((unsigned)(r >> 48) >= thr ? 8u : 0u);
}

__global__ void __launch_bounds__(kThreads)
bias_relu_dropout_kernel(const float* __restrict__ h, const float* __restrict__ bias, float* __restrict__ out,
                         long long n4, int c4, unsigned long long seed, unsigned thr, float scale) {
  ddf::pdl_sync();
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n4) return;
  float4 v = ldg4(h + i * 4);
  if (bias) {
    const float4 b = ldg4(bias + (i % c4) * 4);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  const unsigned k = thr ? keep4(seed, (unsigned long long)i, thr) : 15u;
  v.x = (k & 1u) && v.x > 0.f ? v.x * scale : 0.f;
  v.y = (k & 2u) && v.y > 0.f ? v.y * scale : 0.f;
  v.z = (k & 4u) && v.z > 0.f ? v.z * scale : 0.f;
  v.w = (k & 8u) && v.w > 0.f ? v.w * scale : 0.f;
  *reinterpret_cast<float4*>(out + i * 4) = v;
}

__global__ void __launch_bounds__(kThreads)
relu_dropout_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ out, float* __restrict__ gx,
                        long long n4, float scale) {
  ddf::pdl_sync();
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n4) return;
  const float4 g = ldg4(gy + i * 4), o = ldg4(out + i * 4);
  float4 r;
  r.x = o.x != 0.f ? g.x * scale : 0.f;
  r.y = o.y != 0.f ? g.y * scale : 0.f;
  r.z = o.z != 0.f ? g.z * scale : 0.f;
  r.w = o.w != 0.f ? g.w * scale : 0.f;
  *reinterpret_cast<float4*>(gx + i * 4) = r;
}

// Same as relu_dropout_bwd_kernel plus the bias gradient (column sums of grad_h): a thread keeps its 4
// columns and walks the rows grid-stride, so the [tokens, d_ffn] gradient is not read a second time by
// a separate reduction.  Requires C / 4 <= 256 and 256 % (C / 4) == 0.
__global__ void __launch_bounds__(kThreads)
relu_dropout_bwd_bias_kernel(const float* __restrict__ gy, const float* __restrict__ out, float* __restrict__ gx,
                             float* __restrict__ gbias, long long rows, int c4, float scale) {
  ddf::pdl_sync();
  const int col = threadIdx.x % c4, rl = threadIdx.x / c4, rpb = kThreads / c4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = (long long)blockIdx.x * rpb + rl; r < rows; r += (long long)gridDim.x * rpb) {
    const long long i = r * c4 + col;
    const float4 g = ldg4(gy + i * 4), o = ldg4(out + i * 4);
    float4 v;
    v.x = o.x != 0.f ? g.x * scale : 0.f;
    v.y = o.y != 0.f ? g.y * scale : 0.f;
    v.z = o.z != 0.f ? g.z * scale : 0.f;
    v.w = o.w != 0.f ? g.w * scale : 0.f;
    *reinterpret_cast<float4*>(gx + i * 4) = v;
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  __shared__ float4 sh[kThreads];
  sh[threadIdx.x] = acc;
  __syncthreads();
  if (rl == 0) {
    for (int j = 1; j < rpb; ++j) {
      const float4 t = sh[j * c4 + col];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    atomicAdd(gbias + col * 4, acc.x);
    atomicAdd(gbias + col * 4 + 1, acc.y);
    atomicAdd(gbias + col * 4 + 2, acc.z);
    atomicAdd(gbias + col * 4 + 3, acc.w);
  }
}

// ---- column sums of a tall [rows, C] matrix: the bias gradient of a Linear over all tokens -------------------
// autograd's grad.sum(0) on a [146 k, 128] tensor runs 14x below the HBM rate in ATen (168 us for 75 MB, torch
// profiler, B200); here a thread owns 4 columns and walks the rows grid-stride, one shared-memory fold and one
// atomicAdd per column and CTA.  out must be zeroed by the caller.  C / 4 <= 256 and 256 % (C / 4) == 0.
__global__ void __launch_bounds__(kThreads)
col_sum_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows, int c4) {
  ddf::pdl_sync();
  const int col = threadIdx.x % c4, rl = threadIdx.x / c4, rpb = kThreads / c4;   // threads past rpb * c4 idle
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long step = (long long)gridDim.x * rpb;
  long long r = (long long)blockIdx.x * rpb + rl;
  if (rl < rpb) {
    const float* px = x + col * 4;
    for (; r + 7 * step < rows; r += 8 * step) {   // eight independent 16-byte loads in flight per thread
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = ldg4(px + (r + u * step) * c4 * 4);
#pragma unroll
      for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    for (; r < rows; r += step) {
      const float4 v = ldg4(px + r * c4 * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  __shared__ float4 sh[kThreads];
  sh[threadIdx.x] = acc;
  __syncthreads();
  if (rl == 0) {
    for (int j = 1; j < rpb; ++j) {
      const float4 t = sh[j * c4 + col];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    atomicAdd(out + col * 4, acc.x);
    atomicAdd(out + col * 4 + 1, acc.y);
    atomicAdd(out + col * 4 + 2, acc.z);
    atomicAdd(out + col * 4 + 3, acc.w);
  }
}

// ---- s = a + dropout(b);  y = LayerNorm(s) * gamma + beta.  One warp per row, C = 32 * VPL * 4 ---------
// LPR = lanes per row (32; 16 / 8 for C = 64 / 32: a warp then owns 2 / 4 rows), VPL = float4 vectors per lane
template <int VPL, int LPR>   // C = 4 * LPR * VPL
__global__ void __launch_bounds__(kThreads)
add_dropout_ln_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                          const float* __restrict__ gamma, const float* __restrict__ beta,
                          float* __restrict__ s_out, float* __restrict__ y, float* __restrict__ mean_out,
                          float* __restrict__ rstd_out, long long rows, unsigned long long seed, unsigned thr,
                          float scale, float eps) {
  ddf::pdl_sync();
  constexpr int C = 4 * LPR * VPL;
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const long long wrow = (((long long)blockIdx.x * kThreads + threadIdx.x) >> 5) * (32 / LPR);
  if (wrow >= rows) return;                      // warp-uniform
  const bool valid = wrow + lane / LPR < rows;   // a warp's last rows may not exist: computed on a copy, not stored
  const long long row = valid ? wrow + lane / LPR : rows - 1;
  float4 v[VPL];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const long long e4 = row * (C / 4) + j * LPR + sub;
    v[j] = ldg4(a + e4 * 4);
    if (b) {
      float4 d = ldg4(b + e4 * 4);
      const unsigned k = thr ? keep4(seed, (unsigned long long)e4, thr) : 15u;
      v[j].x += (k & 1u) ? d.x * scale : 0.f;
      v[j].y += (k & 2u) ? d.y * scale : 0.f;
      v[j].z += (k & 4u) ? d.z * scale : 0.f;
      v[j].w += (k & 8u) ? d.w * scale : 0.f;
    }
    sum += v[j].x + v[j].y + v[j].z + v[j].w;
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * (1.f / C);
  float var = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const float dx = v[j].x - mean, dy = v[j].y - mean, dz = v[j].z - mean, dw = v[j].w - mean;
    var += dx * dx + dy * dy + dz * dz + dw * dw;
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var * (1.f / C) + eps);
  if (!valid) return;
  if (sub == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const long long e4 = row * (C / 4) + j * LPR + sub;
    const int c = (j * LPR + sub) * 4;
    if (s_out) *reinterpret_cast<float4*>(s_out + e4 * 4) = v[j];
    const float4 g = ldg4(gamma + c), be = ldg4(beta + c);
    float4 o;
    o.x = (v[j].x - mean) * rstd * g.x + be.x;
    o.y = (v[j].y - mean) * rstd * g.y + be.y;
    o.z = (v[j].z - mean) * rstd * g.z + be.z;
    o.w = (v[j].w - mean) * rstd * g.w + be.w;
    *reinterpret_cast<float4*>(y + e4 * 4) = o;
  }
}

// gs = rstd * (gy*gamma - mean_c(gy*gamma) - xhat * mean_c(gy*gamma*xhat));  ga = gs;  gb = gs * keep * scale
// ggamma += sum_rows gy * xhat, gbeta += sum_rows gy  (per-CTA partials in shared memory, then atomics)
template <int VPL, int LPR>
__global__ void __launch_bounds__(kThreads)
add_dropout_ln_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ s,
                          const float* __restrict__ gamma, const float* __restrict__ mean_in,
                          const float* __restrict__ rstd_in, float* __restrict__ ga, float* __restrict__ gb,
                          float* __restrict__ ggamma, float* __restrict__ gbeta, long long rows,
                          unsigned long long seed, unsigned thr, float scale) {
  ddf::pdl_sync();
  constexpr int C = 4 * LPR * VPL, RPW = 32 / LPR;
  __shared__ float sh_g[C], sh_b[C];
  for (int i = threadIdx.x; i < C; i += kThreads) sh_g[i] = sh_b[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const long long warps = ((long long)gridDim.x * kThreads) >> 5;
  float4 acc_g[VPL], acc_b[VPL];
#pragma unroll
  for (int j = 0; j < VPL; ++j) acc_g[j] = acc_b[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long wr = (((long long)blockIdx.x * kThreads + threadIdx.x) >> 5) * RPW; wr < rows; wr += warps * RPW) {
    const bool valid = wr + lane / LPR < rows;
    const long long row = valid ? wr + lane / LPR : rows - 1;
    const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
    float4 g[VPL], xh[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const long long e4 = row * (C / 4) + j * LPR + sub;
      float4 gyv = ldg4(gy + e4 * 4);
      if (!valid) gyv = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 sv = ldg4(s + e4 * 4), gm = ldg4(gamma + (j * LPR + sub) * 4);
      xh[j] = make_float4((sv.x - mean) * rstd, (sv.y - mean) * rstd, (sv.z - mean) * rstd, (sv.w - mean) * rstd);
      acc_b[j].x += gyv.x; acc_b[j].y += gyv.y; acc_b[j].z += gyv.z; acc_b[j].w += gyv.w;
      acc_g[j].x += gyv.x * xh[j].x; acc_g[j].y += gyv.y * xh[j].y;
      acc_g[j].z += gyv.z * xh[j].z; acc_g[j].w += gyv.w * xh[j].w;
      g[j] = make_float4(gyv.x * gm.x, gyv.y * gm.y, gyv.z * gm.z, gyv.w * gm.w);
      s1 += g[j].x + g[j].y + g[j].z + g[j].w;
      s2 += g[j].x * xh[j].x + g[j].y * xh[j].y + g[j].z * xh[j].z + g[j].w * xh[j].w;
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 * (1.f / C), m2 = s2 * (1.f / C);
    if (!valid) continue;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const long long e4 = row * (C / 4) + j * LPR + sub;
      float4 r;
      r.x = rstd * (g[j].x - m1 - xh[j].x * m2);
      r.y = rstd * (g[j].y - m1 - xh[j].y * m2);
      r.z = rstd * (g[j].z - m1 - xh[j].z * m2);
      r.w = rstd * (g[j].w - m1 - xh[j].w * m2);
      if (ga) *reinterpret_cast<float4*>(ga + e4 * 4) = r;
      if (gb) {
        const unsigned k = thr ? keep4(seed, (unsigned long long)e4, thr) : 15u;
        float4 d;
        d.x = (k & 1u) ? r.x * scale : 0.f;
        d.y = (k & 2u) ? r.y * scale : 0.f;
        d.z = (k & 4u) ? r.z * scale : 0.f;
        d.w = (k & 8u) ? r.w * scale : 0.f;
        *reinterpret_cast<float4*>(gb + e4 * 4) = d;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int c = (j * LPR + sub) * 4;
    atomicAdd(&sh_g[c], acc_g[j].x); atomicAdd(&sh_g[c + 1], acc_g[j].y);
    atomicAdd(&sh_g[c + 2], acc_g[j].z); atomicAdd(&sh_g[c + 3], acc_g[j].w);
    atomicAdd(&sh_b[c], acc_b[j].x); atomicAdd(&sh_b[c + 1], acc_b[j].y);
    atomicAdd(&sh_b[c + 2], acc_b[j].z); atomicAdd(&sh_b[c + 3], acc_b[j].w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += kThreads) {
    if (ggamma) atomicAdd(ggamma + i, sh_g[i]);
    if (gbeta) atomicAdd(gbeta + i, sh_b[i]);
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline unsigned threshold(float p) {
  if (p <= 0.f) return 0u;
  const float t = p * 65536.f + 0.5f;
  return t >= 65535.f ? 65535u : (unsigned)t;
}

}  // namespace

// out[n, C] = dropout(relu(h + bias)); out may alias h. p = drop probability (0 in eval), the kept
// values are scaled by 1/(1-p). C % 4 == 0, 16-byte aligned pointers. bias may be NULL.
extern "C" int ddf_bias_relu_dropout_forward(const float* h, const float* bias, float* out, int64_t n,
                                             int64_t C, float p, uint64_t seed, void* stream_) {
  DDF_CHECK_ARG(n >= 0 && C > 0 && C % 4 == 0 && p >= 0.f && p < 1.f, "bias_relu_dropout: bad arguments");
  if (n == 0) return DDF_OK;
  DDF_CHECK_ARG(h && out && aligned16(h) && aligned16(out) && aligned16(bias), "bias_relu_dropout: null or misaligned pointer");
  const long long n4 = n * C / 4;
  DDF_LAUNCH(bias_relu_dropout_kernel, (unsigned)ddf::cdiv(n4, kThreads), kThreads, 0, (cudaStream_t)stream_, h, bias,
             out, n4, (int)(C / 4), (unsigned long long)seed, threshold(p), 1.f / (1.f - p));
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// grad_h = grad_out * (out != 0) / (1 - p): `out` is the forward result (no mask tensor is kept).
// grad_bias [C] (optional) is ACCUMULATED into (the caller zeroes it): column sums of grad_h.
extern "C" int ddf_bias_relu_dropout_backward(const float* grad_out, const float* out, float* grad_h,
                                              float* grad_bias, int64_t n, int64_t C, float p, void* stream_) {
  DDF_CHECK_ARG(n >= 0 && C > 0 && C % 4 == 0 && p >= 0.f && p < 1.f, "bias_relu_dropout_backward: bad arguments");
  if (n == 0) return DDF_OK;
  DDF_CHECK_ARG(grad_out && out && grad_h && aligned16(grad_out) && aligned16(out) && aligned16(grad_h),
                "bias_relu_dropout_backward: null or misaligned pointer");
  const int c4 = (int)(C / 4);
  if (grad_bias && c4 <= kThreads && kThreads % c4 == 0) {
    const int rpb = kThreads / c4;
    long long grid = ddf::cdiv(n, rpb);
    if (grid > 8 * ddf::kNumSM) grid = 8 * ddf::kNumSM;
    DDF_LAUNCH_PDL(relu_dropout_bwd_bias_kernel, (unsigned)grid, kThreads, 0, (cudaStream_t)stream_, grad_out, out, grad_h,
               grad_bias, (long long)n, c4, 1.f / (1.f - p));
  } else {
    DDF_CHECK_ARG(grad_bias == nullptr, "bias_relu_dropout_backward: grad_bias needs C/4 to divide 256");
    DDF_LAUNCH(relu_dropout_bwd_kernel, (unsigned)ddf::cdiv(n * c4, kThreads), kThreads, 0, (cudaStream_t)stream_,
               grad_out, out, grad_h, (long long)(n * c4), 1.f / (1.f - p));
  }
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

#define DDF_LN_DISPATCH(C, CALL)                                          \
  switch (C) {                                                            \
    case 32: { constexpr int VPL = 1, LPR = 8; CALL; } break;             \
    case 64: { constexpr int VPL = 1, LPR = 16; CALL; } break;            \
    case 128: { constexpr int VPL = 1, LPR = 32; CALL; } break;           \
    case 256: { constexpr int VPL = 2, LPR = 32; CALL; } break;           \
    case 512: { constexpr int VPL = 4, LPR = 32; CALL; } break;           \
    default:                                                              \
      ddf::set_error("add_dropout_layer_norm: C must be 32, 64, 128, 256 or 512, got %lld", (long long)(C)); \
      return DDF_ERR_ARG;                                                 \
  }

// s = a + dropout(b) (b may be NULL: s = a);  y = LayerNorm(s) * gamma + beta over the last dim C.
// s_out (optional) receives s, mean / rstd [rows] are saved for backward.
extern "C" int ddf_add_dropout_layer_norm_forward(const float* a, const float* b, const float* gamma,
                                                  const float* beta, float* s_out, float* y, float* mean,
                                                  float* rstd, int64_t rows, int64_t C, float p,
                                                  uint64_t seed, float eps, void* stream_) {
  DDF_CHECK_ARG(rows >= 0 && p >= 0.f && p < 1.f, "add_dropout_layer_norm: bad arguments");
  if (rows == 0) return DDF_OK;
  DDF_CHECK_ARG(a && gamma && beta && y && mean && rstd, "add_dropout_layer_norm: null pointer");
  DDF_CHECK_ARG(aligned16(a) && aligned16(b) && aligned16(gamma) && aligned16(beta) && aligned16(s_out) && aligned16(y),
                "add_dropout_layer_norm: misaligned pointer");
  const long long wrows = C >= 128 ? rows : ddf::cdiv(rows, 128 / C);     // rows of warps
  const unsigned grid = (unsigned)ddf::cdiv(wrows * 32, kThreads);
  DDF_LN_DISPATCH(C, DDF_LAUNCH_PDL((add_dropout_ln_fwd_kernel<VPL, LPR>), grid, kThreads, 0, (cudaStream_t)stream_, a, b, gamma,
                                beta, s_out, y, mean, rstd, (long long)rows, (unsigned long long)seed, threshold(p),
                                1.f / (1.f - p), eps));
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// grad_gamma / grad_beta [C] are ACCUMULATED into (caller zeroes them); grad_a / grad_b may be NULL.
extern "C" int ddf_add_dropout_layer_norm_backward(const float* grad_y, const float* s, const float* gamma,
                                                   const float* mean, const float* rstd, float* grad_a,
                                                   float* grad_b, float* grad_gamma, float* grad_beta,
                                                   int64_t rows, int64_t C, float p, uint64_t seed,
                                                   void* stream_) {
  DDF_CHECK_ARG(rows >= 0 && p >= 0.f && p < 1.f, "add_dropout_layer_norm_backward: bad arguments");
  if (rows == 0) return DDF_OK;
  DDF_CHECK_ARG(grad_y && s && gamma && mean && rstd, "add_dropout_layer_norm_backward: null pointer");
  long long grid = ddf::cdiv((C >= 128 ? rows : ddf::cdiv(rows, 128 / C)) * 32, kThreads);
  if (grid > 4 * ddf::kNumSM) grid = 4 * ddf::kNumSM;   // grid-stride: bounded number of atomics on grad_gamma / beta
  DDF_LN_DISPATCH(C, DDF_LAUNCH_PDL((add_dropout_ln_bwd_kernel<VPL, LPR>), (unsigned)grid, kThreads, 0, (cudaStream_t)stream_, grad_y,
                                s, gamma, mean, rstd, grad_a, grad_b, grad_gamma, grad_beta, (long long)rows,
                                (unsigned long long)seed, threshold(p), 1.f / (1.f - p)));
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// out [C] = sum over the rows of x [rows, C] (zeroed inside).  C % 4 == 0, C / 4 <= 256, 256 % (C / 4) == 0.
extern "C" int ddf_col_sum(const float* x, float* out, int64_t rows, int64_t C, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(rows >= 0 && C > 0 && C % 4 == 0 && C / 4 <= kThreads, "col_sum: C must be a multiple of 4, <= %d (C=%lld)",
                4 * kThreads, (long long)C);
  DDF_CHECK_ARG(out != nullptr, "col_sum: null out");
  DDF_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)C, stream));
  if (rows == 0) return DDF_OK;
  DDF_CHECK_ARG(x != nullptr && aligned16(x), "col_sum: null / misaligned x");
  const int c4 = (int)(C / 4), rpb = kThreads / c4;
  long long grid = ddf::cdiv(rows, rpb);
  if (grid > 4 * ddf::kNumSM) grid = 4 * ddf::kNumSM;
  DDF_LAUNCH_PDL(col_sum_kernel, (unsigned)grid, kThreads, 0, stream, x, out, (long long)rows, c4);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// ---- bi-directional gated fusion of the two query streams (<proj>/models/model_utils/attentions.py:89-117) -------
//   u = fuse_in ? f1 + f2 : (f1 for s1, f2 for s2);  s1 = sigmoid(wb . u + bb), s2 = sigmoid(wa . u + ba)
//   o1 = f1 + f2 * s1,  o2 = f2 + f1 * s2                      (BiGateSum1D / BiGateSum1D_2)
// The module chain is nine elementwise / reduction launches over [tokens, C] forward and about twice that backward;
// here one pass each way: a warp owns a row (C = 128 * VPL), the two dot products are shuffle reductions, the
// gates [rows, 2] are kept for backward.
namespace {
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

template <int VPL>
__global__ void __launch_bounds__(kThreads)
bigate_sum_fwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ wb,
                      const float* __restrict__ bb, const float* __restrict__ wa, const float* __restrict__ ba,
                      float* __restrict__ o1, float* __restrict__ o2, float* __restrict__ gates, long long rows,
                      int fuse_in) {
  ddf::pdl_sync();
  constexpr int C = 128 * VPL;
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
  if (row >= rows) return;
  float4 a[VPL], b[VPL];
  float d1 = 0.f, d2 = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = (v * 32 + lane) * 4;
    a[v] = ldg4(f1 + row * C + c);
    b[v] = ldg4(f2 + row * C + c);
    const float4 w1 = ldg4(wb + c), w2 = ldg4(wa + c);
    if (fuse_in) {
      const float4 u = make_float4(a[v].x + b[v].x, a[v].y + b[v].y, a[v].z + b[v].z, a[v].w + b[v].w);
      d1 += dot4(w1, u);
      d2 += dot4(w2, u);
    } else {
      d1 += dot4(w1, a[v]);
      d2 += dot4(w2, b[v]);
    }
  }
  const float s1 = 1.f / (1.f + expf(-(warp_sum(d1) + (bb ? __ldg(bb) : 0.f))));
  const float s2 = 1.f / (1.f + expf(-(warp_sum(d2) + (ba ? __ldg(ba) : 0.f))));
  if (lane == 0) *reinterpret_cast<float2*>(gates + row * 2) = make_float2(s1, s2);
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = (v * 32 + lane) * 4;
    *reinterpret_cast<float4*>(o1 + row * C + c) =
        make_float4(fmaf(b[v].x, s1, a[v].x), fmaf(b[v].y, s1, a[v].y), fmaf(b[v].z, s1, a[v].z), fmaf(b[v].w, s1, a[v].w));
    *reinterpret_cast<float4*>(o2 + row * C + c) =
        make_float4(fmaf(a[v].x, s2, b[v].x), fmaf(a[v].y, s2, b[v].y), fmaf(a[v].z, s2, b[v].z), fmaf(a[v].w, s2, b[v].w));
  }
}

// go1 / go2 may be NULL (an output that never reaches the loss).  gwb / gwa [C], gbb / gba [1] are accumulated
// (zeroed by the caller): per-lane partial sums over the warp's rows, one shared-memory fold per CTA, atomics.
template <int VPL>
__global__ void __launch_bounds__(kThreads)
bigate_sum_bwd_kernel(const float* __restrict__ go1, const float* __restrict__ go2, const float* __restrict__ f1,
                      const float* __restrict__ f2, const float* __restrict__ gates, const float* __restrict__ wb,
                      const float* __restrict__ wa, float* __restrict__ gf1, float* __restrict__ gf2,
                      float* __restrict__ gwb, float* __restrict__ gbb, float* __restrict__ gwa,
                      float* __restrict__ gba, long long rows, int fuse_in) {
  ddf::pdl_sync();
  constexpr int C = 128 * VPL;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 w1[VPL], w2[VPL], aw1[VPL], aw2[VPL];
  float ab1 = 0.f, ab2 = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = (v * 32 + lane) * 4;
    w1[v] = ldg4(wb + c);
    w2[v] = ldg4(wa + c);
    aw1[v] = aw2[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const long long wstride = (long long)gridDim.x * (kThreads / 32);
  for (long long row = (long long)blockIdx.x * (kThreads / 32) + warp; row < rows; row += wstride) {
    float4 a[VPL], b[VPL], g1[VPL], g2[VPL];
    float e1 = 0.f, e2 = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const long long o = row * C + (v * 32 + lane) * 4;
      a[v] = ldg4(f1 + o);
      b[v] = ldg4(f2 + o);
      g1[v] = go1 ? ldg4(go1 + o) : make_float4(0.f, 0.f, 0.f, 0.f);
      g2[v] = go2 ? ldg4(go2 + o) : make_float4(0.f, 0.f, 0.f, 0.f);
      e1 += dot4(g1[v], b[v]);     // d o1 / d s1 = f2
      e2 += dot4(g2[v], a[v]);     // d o2 / d s2 = f1
    }
    const float2 s = __ldg(reinterpret_cast<const float2*>(gates) + row);
    const float z1 = warp_sum(e1) * s.x * (1.f - s.x), z2 = warp_sum(e2) * s.y * (1.f - s.y);
    ab1 += z1;
    ab2 += z2;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const long long o = row * C + (v * 32 + lane) * 4;
      // gradient through the gate inputs
      const float4 u1 = make_float4(z1 * w1[v].x, z1 * w1[v].y, z1 * w1[v].z, z1 * w1[v].w);
      const float4 u2 = make_float4(z2 * w2[v].x, z2 * w2[v].y, z2 * w2[v].z, z2 * w2[v].w);
      float4 r1, r2;
      r1.x = fmaf(g2[v].x, s.y, g1[v].x); r1.y = fmaf(g2[v].y, s.y, g1[v].y);
      r1.z = fmaf(g2[v].z, s.y, g1[v].z); r1.w = fmaf(g2[v].w, s.y, g1[v].w);
      r2.x = fmaf(g1[v].x, s.x, g2[v].x); r2.y = fmaf(g1[v].y, s.x, g2[v].y);
      r2.z = fmaf(g1[v].z, s.x, g2[v].z); r2.w = fmaf(g1[v].w, s.x, g2[v].w);
      float4 in1, in2;           // the vectors the two gates were computed from
      if (fuse_in) {
        const float4 t = make_float4(u1.x + u2.x, u1.y + u2.y, u1.z + u2.z, u1.w + u2.w);
        r1.x += t.x; r1.y += t.y; r1.z += t.z; r1.w += t.w;
        r2.x += t.x; r2.y += t.y; r2.z += t.z; r2.w += t.w;
        in1 = in2 = make_float4(a[v].x + b[v].x, a[v].y + b[v].y, a[v].z + b[v].z, a[v].w + b[v].w);
      } else {
        r1.x += u1.x; r1.y += u1.y; r1.z += u1.z; r1.w += u1.w;
        r2.x += u2.x; r2.y += u2.y; r2.z += u2.z; r2.w += u2.w;
        in1 = a[v];
        in2 = b[v];
      }
      if (gf1) *reinterpret_cast<float4*>(gf1 + o) = r1;
      if (gf2) *reinterpret_cast<float4*>(gf2 + o) = r2;
      aw1[v].x = fmaf(z1, in1.x, aw1[v].x); aw1[v].y = fmaf(z1, in1.y, aw1[v].y);
      aw1[v].z = fmaf(z1, in1.z, aw1[v].z); aw1[v].w = fmaf(z1, in1.w, aw1[v].w);
      aw2[v].x = fmaf(z2, in2.x, aw2[v].x); aw2[v].y = fmaf(z2, in2.y, aw2[v].y);
      aw2[v].z = fmaf(z2, in2.z, aw2[v].z); aw2[v].w = fmaf(z2, in2.w, aw2[v].w);
    }
  }
  // fold the 8 warps of the CTA, then one atomic per column and CTA
  __shared__ float4 sh[2][kThreads / 32][32 * VPL];
  __shared__ float shb[2][kThreads / 32];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    sh[0][warp][v * 32 + lane] = aw1[v];
    sh[1][warp][v * 32 + lane] = aw2[v];
  }
  if (lane == 0) {
    shb[0][warp] = ab1;
    shb[1][warp] = ab2;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * 32 * VPL; i += kThreads) {
    const int which = i / (32 * VPL), col = i % (32 * VPL);
    float4 t = sh[which][0][col];
    for (int w = 1; w < kThreads / 32; ++w) {
      const float4 x = sh[which][w][col];
      t.x += x.x; t.y += x.y; t.z += x.z; t.w += x.w;
    }
    float* dst = (which ? gwa : gwb);
    if (dst) {
      atomicAdd(dst + col * 4, t.x);
      atomicAdd(dst + col * 4 + 1, t.y);
      atomicAdd(dst + col * 4 + 2, t.z);
      atomicAdd(dst + col * 4 + 3, t.w);
    }
  }
  if (threadIdx.x < 2) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += shb[threadIdx.x][w];
    float* dst = threadIdx.x ? gba : gbb;
    if (dst) atomicAdd(dst, t);
  }
}
}  // namespace

#define DDF_GATE_DISPATCH(C, CALL)                    \
  switch (C) {                                        \
    case 128: { constexpr int VPL = 1; CALL; } break; \
    case 256: { constexpr int VPL = 2; CALL; } break; \
    default:                                          \
      ddf::set_error("bigate_sum: C must be 128 or 256, got %lld", (long long)(C)); \
      return DDF_ERR_ARG;                             \
  }

// o1 = f1 + f2 * s1, o2 = f2 + f1 * s2 with s1 = sigmoid(wb . u1 + bb), s2 = sigmoid(wa . u2 + ba); u1 = u2 = f1 + f2
// when fuse_in (BiGateSum1D_2) else u1 = f1, u2 = f2 (BiGateSum1D).  gates [rows, 2] = (s1, s2) for backward.
extern "C" int ddf_bigate_sum_forward(const float* f1, const float* f2, const float* wb, const float* bb,
                                      const float* wa, const float* ba, float* o1, float* o2, float* gates,
                                      int64_t rows, int64_t C, int fuse_in, void* stream_) {
  DDF_CHECK_ARG(rows >= 0, "bigate_sum_forward: bad rows");
  if (rows == 0) return DDF_OK;
  DDF_CHECK_ARG(f1 && f2 && wb && wa && o1 && o2 && gates, "bigate_sum_forward: null pointer");
  DDF_CHECK_ARG(aligned16(f1) && aligned16(f2) && aligned16(wb) && aligned16(wa) && aligned16(o1) && aligned16(o2) &&
                    (reinterpret_cast<uintptr_t>(gates) & 7u) == 0,
                "bigate_sum_forward: misaligned pointer");
  const unsigned grid = (unsigned)ddf::cdiv(rows * 32, kThreads);
  DDF_GATE_DISPATCH(C, DDF_LAUNCH_PDL(bigate_sum_fwd_kernel<VPL>, grid, kThreads, 0, (cudaStream_t)stream_, f1, f2, wb, bb,
                                  wa, ba, o1, o2, gates, (long long)rows, fuse_in));
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// grad_o1 / grad_o2 may be NULL (treated as zero); grad_f1 / grad_f2 may be NULL; grad_wb / grad_wa [C] and
// grad_bb / grad_ba [1] are overwritten (NULL = not wanted).
extern "C" int ddf_bigate_sum_backward(const float* grad_o1, const float* grad_o2, const float* f1, const float* f2,
                                       const float* gates, const float* wb, const float* wa, float* grad_f1,
                                       float* grad_f2, float* grad_wb, float* grad_bb, float* grad_wa,
                                       float* grad_ba, int64_t rows, int64_t C, int fuse_in, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(rows >= 0 && (C == 128 || C == 256), "bigate_sum_backward: C must be 128 or 256");
  if (grad_wb) DDF_CUDA(cudaMemsetAsync(grad_wb, 0, sizeof(float) * (size_t)C, stream));
  if (grad_wa) DDF_CUDA(cudaMemsetAsync(grad_wa, 0, sizeof(float) * (size_t)C, stream));
  if (grad_bb) DDF_CUDA(cudaMemsetAsync(grad_bb, 0, sizeof(float), stream));
  if (grad_ba) DDF_CUDA(cudaMemsetAsync(grad_ba, 0, sizeof(float), stream));
  if (rows == 0) return DDF_OK;
  DDF_CHECK_ARG(f1 && f2 && gates && wb && wa, "bigate_sum_backward: null pointer");
  DDF_CHECK_ARG(aligned16(grad_o1) && aligned16(grad_o2) && aligned16(f1) && aligned16(f2) && aligned16(grad_f1) &&
                    aligned16(grad_f2) && aligned16(wb) && aligned16(wa),
                "bigate_sum_backward: misaligned pointer");
  long long grid = ddf::cdiv(rows * 32, kThreads);
  if (grid > 4 * ddf::kNumSM) grid = 4 * ddf::kNumSM;
  DDF_GATE_DISPATCH(C, DDF_LAUNCH_PDL(bigate_sum_bwd_kernel<VPL>, (unsigned)grid, kThreads, 0, stream, grad_o1, grad_o2, f1,
                                  f2, gates, wb, wa, grad_f1, grad_f2, grad_wb, grad_bb, grad_wa, grad_ba,
                                  (long long)rows, fuse_in));
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}
