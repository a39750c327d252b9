// BatchNorm1d over sparse-voxel features [N, C] fused with the residual add and ReLU that follow it
// in every block of the sparse encoders, forward and backward, for sm_100a.
//
// Replaces, per conv layer of the reference (TransFusion/mmdet3d/ops/sparse_block.py:102-120, 153-185:
// conv -> BN1d(eps 1e-3, momentum 0.01) -> [+ identity] -> ReLU), the chain of separate elementwise
// kernels over (N, C): batch_norm statistics + transform, add, clamp in forward (4 launches) and
// threshold_backward, batch_norm backward reduce + elemt, add in backward (4-5 launches) by
//   forward : bn_stats_kernel (column sums; the partials are folded in two levels by the last CTAs to finish, the
//             very last one writes mean / invstd and updates the running statistics) + bn_apply_kernel
//   backward: bn_bwd_reduce_kernel (ReLU mask applied on the fly; last CTA writes grad_weight /
//             grad_bias and the two per-channel coefficients) + bn_bwd_apply_kernel
// HBM-bound streaming kernels: rows are read as 16-byte vectors, a thread owns 4 channels of a row,
// per-thread accumulation in fp64 (so E[x^2] - mean^2 is safe), deterministic (no float atomics:
// fixed-order fold of per-CTA partials).
// Algorithmic bytes: forward 4*N*C*(2 reads + [1 residual read] + 1 write); backward
// 4*N*C*(3 reads + 3 reads + 1-2 writes).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxGrid = 4 * ddf::kNumSM;

struct Acc8 {
  double v[8];
};

// Two-level, fixed-order fold of the per-CTA partials.  One CTA folding all 592 partials was the tail of every
// statistics launch: 2 * C columns x 592 doubles (1.2 MB at C = 128) through ONE SM, 11-22 us of a 30-50 us kernel.
// Now the last CTA to arrive in each group of kFoldGroup consecutive CTAs folds that group (16 loads per column, all in
// flight, on 37 SMs at once) and the last group to finish folds the 37 group sums.  Still deterministic: the order of
// the additions depends on the indices only, not on the arrival order.
constexpr int kFoldGroup = 16;
constexpr int kMaxGroups = (kMaxGrid + kFoldGroup - 1) / kFoldGroup;

// counters[0]: groups finished; counters[1 + g]: CTAs of group g finished.  All zero between launches.
__device__ __forceinline__ bool arrive_last(unsigned* counter, unsigned expected) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(counter, 1u);
    s_last = ticket == expected - 1;
    if (s_last) *counter = 0;  // ready for the next launch on this workspace
  }
  __syncthreads();
  const bool last = s_last;
  if (last) __threadfence();
  __syncthreads();             // s_last is reused by the second level
  return last;
}

// Block-level fold of the per-thread (sum[4], sq[4]) accumulators over the rows of the block, one partial per CTA
// (part[blockIdx.x][0..1][C]), then the two-level fold.  Returns true in the ONE CTA that ends up with the totals of
// all 2 * C columns in tot[] (shared memory).
__device__ __forceinline__ bool fold_and_publish(Acc8& a, int tpr, int cg, int rl, int rpb, int C,
                                                 double* __restrict__ part, double* __restrict__ part2,
                                                 unsigned* counters, double* tot) {
  __shared__ double sm[kThreads * 8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[i * kThreads + threadIdx.x] = a.v[i];
  __syncthreads();
  for (int off = rpb >> 1; off > 0; off >>= 1) {
    if (rl < off) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sm[i * kThreads + threadIdx.x] += sm[i * kThreads + threadIdx.x + off * tpr];
    }
    __syncthreads();
  }
  const int cols = 2 * C;
  if (rl == 0) {
    double* p = part + (long long)blockIdx.x * cols + cg * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      p[i] = sm[i * kThreads + threadIdx.x];
      p[C + i] = sm[(4 + i) * kThreads + threadIdx.x];
    }
  }
  const int grp = blockIdx.x / kFoldGroup, ngroups = (gridDim.x + kFoldGroup - 1) / kFoldGroup;
  const int g0 = grp * kFoldGroup;
  const int gsize = min(kFoldGroup, (int)gridDim.x - g0);
  if (!arrive_last(counters + 1 + grp, (unsigned)gsize)) return false;
  for (int col = threadIdx.x; col < cols; col += kThreads) {
    const double* p = part + (long long)g0 * cols + col;
    double v[kFoldGroup];
#pragma unroll
    for (int u = 0; u < kFoldGroup; ++u) v[u] = u < gsize ? __ldcg(p + (long long)u * cols) : 0.0;
    double s = v[0];
#pragma unroll
    for (int u = 1; u < kFoldGroup; ++u) s += v[u];
    part2[(long long)grp * cols + col] = s;
  }
  if (!arrive_last(counters, (unsigned)ngroups)) return false;
  for (int col = threadIdx.x; col < cols; col += kThreads) {
    const double* p = part2 + col;
    double s = 0.0;
    int g = 0;
    for (; g + kFoldGroup <= ngroups; g += kFoldGroup) {
      double v[kFoldGroup];
#pragma unroll
      for (int u = 0; u < kFoldGroup; ++u) v[u] = __ldcg(p + (long long)(g + u) * cols);
#pragma unroll
      for (int u = 0; u < kFoldGroup; ++u) s += v[u];
    }
    for (; g < ngroups; ++g) s += __ldcg(p + (long long)g * cols);
    tot[col] = s;
  }
  __syncthreads();
  return true;
}

__global__ void __launch_bounds__(kThreads)
bn_stats_kernel(const float* __restrict__ x, int n, int C, double* __restrict__ part, double* __restrict__ part2,
                unsigned* counter, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                float* running_mean, float* running_var, float momentum, float eps) {
  ddf::pdl_sync();
  extern __shared__ double tot[];  // [2*C]
  const int tpr = C >> 2, rpb = kThreads / tpr;
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  Acc8 a;
#pragma unroll
  for (int i = 0; i < 8; ++i) a.v[i] = 0.0;
  auto add = [&](const float4 v) {
    a.v[0] += v.x; a.v[1] += v.y; a.v[2] += v.z; a.v[3] += v.w;
    a.v[4] += (double)v.x * v.x; a.v[5] += (double)v.y * v.y;
    a.v[6] += (double)v.z * v.z; a.v[7] += (double)v.w * v.w;
  };
  const long long step = (long long)gridDim.x * rpb;
  long long r = (long long)blockIdx.x * rpb + rl;
  const float* px = x + cg * 4;
  for (; r + 7 * step < n; r += 8 * step) {   // eight independent 16-byte loads in flight per thread
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = ldg4(px + (r + u * step) * C);
    // the eight rows are folded in fp32 first (relative error 1e-7 per group), one fp64 update per group: the
    // fp32 -> fp64 conversions, not the loads, were the inner-loop cost
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s1.x += v[u].x; s1.y += v[u].y; s1.z += v[u].z; s1.w += v[u].w;
      s2.x = fmaf(v[u].x, v[u].x, s2.x); s2.y = fmaf(v[u].y, v[u].y, s2.y);
      s2.z = fmaf(v[u].z, v[u].z, s2.z); s2.w = fmaf(v[u].w, v[u].w, s2.w);
    }
    a.v[0] += s1.x; a.v[1] += s1.y; a.v[2] += s1.z; a.v[3] += s1.w;
    a.v[4] += s2.x; a.v[5] += s2.y; a.v[6] += s2.z; a.v[7] += s2.w;
  }
  for (; r < n; r += step) add(ldg4(px + r * C));
  if (!fold_and_publish(a, tpr, cg, rl, rpb, C, part, part2, counter, tot)) return;
  for (int c = threadIdx.x; c < C; c += kThreads) {
    const double mean = tot[c] / n;
    double var = tot[C + c] / n - mean * mean;  // biased, used for normalisation
    if (var < 0.0) var = 0.0;
    save_mean[c] = (float)mean;
    save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) {
      const double unbiased = n > 1 ? var * ((double)n / (double)(n - 1)) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
}

// y = relu?((x - mean) * invstd * w + b + res?).  stat_is_var: `invstd` holds a variance (eval mode).
__global__ void __launch_bounds__(kThreads)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ res,
                const float* __restrict__ mean, const float* __restrict__ invstd,
                const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ y,
                long long n4, int C, int relu, int stat_is_var, float eps) {
  ddf::pdl_sync();
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n4) return;
  const int c = (int)(i % (C >> 2)) * 4;
  const float4 v = ldg4(x + i * 4);
  const float4 m = ldg4(mean + c);
  float4 s = ldg4(invstd + c);
  if (stat_is_var) {
    s.x = 1.f / sqrtf(s.x + eps); s.y = 1.f / sqrtf(s.y + eps);
    s.z = 1.f / sqrtf(s.z + eps); s.w = 1.f / sqrtf(s.w + eps);
  }
  if (w) {
    const float4 g = ldg4(w + c);
    s.x *= g.x; s.y *= g.y; s.z *= g.z; s.w *= g.w;
  }
  float4 o = b ? ldg4(b + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  o.x = fmaf(v.x - m.x, s.x, o.x); o.y = fmaf(v.y - m.y, s.y, o.y);
  o.z = fmaf(v.z - m.z, s.z, o.z); o.w = fmaf(v.w - m.w, s.w, o.w);
  if (res) {
    const float4 r = ldg4(res + i * 4);
    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
  }
  if (relu) {
    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
  }
  *reinterpret_cast<float4*>(y + i * 4) = o;
}

// bn_apply for layers whose output feeds a bf16x3 tcgen05 conv: the same pass also writes the conv's operand format
// ([32 x bf16 hi | 32 x bf16 lo] per 32 channels, csrc/sparse_conv.cu ddf_split_bf16x3) and the tf32-rounded copy the
// wgrad kernels read - the separate split pass re-read y for every conv input (31 launches, 1.2 ms of a step).
// One thread = 8 consecutive channels (one 16-byte store of hi, one of lo).  C % 32 == 0.
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__global__ void __launch_bounds__(kThreads)
bn_apply_split_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ mean,
                      const float* __restrict__ invstd, const float* __restrict__ w, const float* __restrict__ b,
                      float* __restrict__ y, uint8_t* __restrict__ split, float* __restrict__ rounded, long long n8,
                      int C, int relu, int stat_is_var, float eps) {
  ddf::pdl_sync();
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (t >= n8) return;
  const long long e = t * 8;
  const long long row = e / C;
  const int c = (int)(e % C);
  float v[8], o[8];
  *reinterpret_cast<float4*>(v) = ldg4(x + e);
  *reinterpret_cast<float4*>(v + 4) = ldg4(x + e + 4);
  float m[8], s[8], bb[8];
  *reinterpret_cast<float4*>(m) = ldg4(mean + c);
  *reinterpret_cast<float4*>(m + 4) = ldg4(mean + c + 4);
  *reinterpret_cast<float4*>(s) = ldg4(invstd + c);
  *reinterpret_cast<float4*>(s + 4) = ldg4(invstd + c + 4);
  if (stat_is_var) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 1.f / sqrtf(s[i] + eps);
  }
  if (w) {
    float g[8];
    *reinterpret_cast<float4*>(g) = ldg4(w + c);
    *reinterpret_cast<float4*>(g + 4) = ldg4(w + c + 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] *= g[i];
  }
  if (b) {
    *reinterpret_cast<float4*>(bb) = ldg4(b + c);
    *reinterpret_cast<float4*>(bb + 4) = ldg4(b + c + 4);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) bb[i] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = fmaf(v[i] - m[i], s[i], bb[i]);
  if (res) {
    float r[8];
    *reinterpret_cast<float4*>(r) = ldg4(res + e);
    *reinterpret_cast<float4*>(r + 4) = ldg4(res + e + 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] += r[i];
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.f);
  }
  *reinterpret_cast<float4*>(y + e) = *reinterpret_cast<float4*>(o);
  *reinterpret_cast<float4*>(y + e + 4) = *reinterpret_cast<float4*>(o + 4);
  unsigned short hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __nv_bfloat16 h = __float2bfloat16_rn(o[i]);
    const __nv_bfloat16 l = __float2bfloat16_rn(o[i] - __bfloat162float(h));
    hi[i] = __bfloat16_as_ushort(h);
    lo[i] = __bfloat16_as_ushort(l);
  }
  uint8_t* dst = split + row * C * 4 + (c >> 5) * 128 + (c & 31) * 2;
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(dst + 64) = *reinterpret_cast<const uint4*>(lo);
  if (rounded) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = tf32_rn(o[i]);
    *reinterpret_cast<float4*>(rounded + e) = *reinterpret_cast<float4*>(o);
    *reinterpret_cast<float4*>(rounded + e + 4) = *reinterpret_cast<float4*>(o + 4);
  }
}

// sums over rows of g and g * xhat, g = gy masked by (y > 0) when relu.
__global__ void __launch_bounds__(kThreads)
bn_bwd_reduce_kernel(const float* __restrict__ gy, const float* __restrict__ y,
                     const float* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ invstd, int n, int C, int relu, int train,
                     double* __restrict__ part, double* __restrict__ part2, unsigned* counter,
                     float* __restrict__ gweight,
                     float* __restrict__ gbias, float* __restrict__ coef) {
  ddf::pdl_sync();
  extern __shared__ double tot[];
  const int tpr = C >> 2, rpb = kThreads / tpr;
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  const float4 m = ldg4(mean + cg * 4), s = ldg4(invstd + cg * 4);
  Acc8 a;
#pragma unroll
  for (int i = 0; i < 8; ++i) a.v[i] = 0.0;
  auto add = [&](float4 g, const float4 yy, const float4 v) {
    if (relu) {
      g.x = yy.x > 0.f ? g.x : 0.f; g.y = yy.y > 0.f ? g.y : 0.f;
      g.z = yy.z > 0.f ? g.z : 0.f; g.w = yy.w > 0.f ? g.w : 0.f;
    }
    a.v[0] += g.x; a.v[1] += g.y; a.v[2] += g.z; a.v[3] += g.w;
    a.v[4] += (double)g.x * ((v.x - m.x) * s.x); a.v[5] += (double)g.y * ((v.y - m.y) * s.y);
    a.v[6] += (double)g.z * ((v.z - m.z) * s.z); a.v[7] += (double)g.w * ((v.w - m.w) * s.w);
  };
  const long long step = (long long)gridDim.x * rpb;
  long long r = (long long)blockIdx.x * rpb + rl;
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
  for (; r + 3 * step < n; r += 4 * step) {   // four rows (twelve 16-byte loads) in flight per thread
    float4 g[4], yy[4], v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long o = (r + u * step) * C + cg * 4;
      g[u] = ldg4(gy + o);
      yy[u] = relu ? ldg4(y + o) : one;
      v[u] = ldg4(x + o);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) add(g[u], yy[u], v[u]);
  }
  for (; r < n; r += step) {
    const long long o = r * C + cg * 4;
    add(ldg4(gy + o), relu ? ldg4(y + o) : one, ldg4(x + o));
  }
  if (!fold_and_publish(a, tpr, cg, rl, rpb, C, part, part2, counter, tot)) return;
  for (int c = threadIdx.x; c < C; c += kThreads) {
    if (gbias) gbias[c] = (float)tot[c];
    if (gweight) gweight[c] = (float)tot[C + c];
    coef[c] = train ? (float)(tot[c] / n) : 0.f;
    coef[C + c] = train ? (float)(tot[C + c] / n) : 0.f;
  }
}

// gx = (g - c1 - xhat * c2) * invstd * w ; gres = g (the gradient of the residual branch).
__global__ void __launch_bounds__(kThreads)
bn_bwd_apply_kernel(const float* __restrict__ gy, const float* __restrict__ y,
                    const float* __restrict__ x, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ w,
                    const float* __restrict__ coef, float* __restrict__ gx,
                    float* __restrict__ gres, long long n4, int C, int relu) {
  ddf::pdl_sync();
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n4) return;
  const int c = (int)(i % (C >> 2)) * 4;
  float4 g = ldg4(gy + i * 4);
  if (relu) {
    const float4 yy = ldg4(y + i * 4);
    g.x = yy.x > 0.f ? g.x : 0.f; g.y = yy.y > 0.f ? g.y : 0.f;
    g.z = yy.z > 0.f ? g.z : 0.f; g.w = yy.w > 0.f ? g.w : 0.f;
  }
  if (gres) *reinterpret_cast<float4*>(gres + i * 4) = g;
  if (!gx) return;
  const float4 v = ldg4(x + i * 4);
  const float4 m = ldg4(mean + c), s = ldg4(invstd + c);
  const float4 c1 = ldg4(coef + c), c2 = ldg4(coef + C + c);
  float4 sw = s;
  if (w) {
    const float4 ww = ldg4(w + c);
    sw.x *= ww.x; sw.y *= ww.y; sw.z *= ww.z; sw.w *= ww.w;
  }
  float4 o;
  o.x = (g.x - c1.x - (v.x - m.x) * s.x * c2.x) * sw.x;
  o.y = (g.y - c1.y - (v.y - m.y) * s.y * c2.y) * sw.y;
  o.z = (g.z - c1.z - (v.z - m.z) * s.z * c2.z) * sw.z;
  o.w = (g.w - c1.w - (v.w - m.w) * s.w * c2.w) * sw.w;
  *reinterpret_cast<float4*>(gx + i * 4) = o;
}

inline bool pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline unsigned stats_grid(int64_t n, int64_t C) {
  const int rpb = kThreads / (int)(C >> 2);
  long long g = ddf::cdiv(n, rpb);
  if (g > kMaxGrid) g = kMaxGrid;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" int64_t ddf_sparse_bn_workspace_bytes(int64_t C) {
  if (!pow2(C) || C < 4 || C > 1024) return -1;
  // per-CTA partials [kMaxGrid][2][C] doubles + group sums [kMaxGroups][2][C] doubles + coefficient block [2][C]
  // floats + arrival counters [1 + kMaxGroups].  Zero-filled by the caller before its FIRST use; the kernels leave the
  // counters at zero.
  return (int64_t)(kMaxGrid + kMaxGroups) * 2 * C * 8 + 2 * C * 4 + 128 + 4 * (1 + kMaxGroups + 25);
}

static inline double* ws_part2(void* ws, int64_t C) {
  return reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + (int64_t)kMaxGrid * 2 * C * 8);
}
static inline float* ws_coef(void* ws, int64_t C) {
  return reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + (int64_t)(kMaxGrid + kMaxGroups) * 2 * C * 8);
}
static inline unsigned* ws_counter(void* ws, int64_t C) {
  return reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) + (int64_t)(kMaxGrid + kMaxGroups) * 2 * C * 8 + 2 * C * 4 +
                                     128);
}

static int bn_forward_impl(const float* x, const float* residual, const float* weight, const float* bias,
                           float* running_mean, float* running_var, float* y, float* save_mean, float* save_invstd,
                           void* split, float* rounded, int64_t n, int64_t C, int training, float momentum, float eps,
                           int relu, void* workspace, void* stream_);

extern "C" int ddf_sparse_bn_forward(const float* x, const float* residual, const float* weight,
                                     const float* bias, float* running_mean, float* running_var,
                                     float* y, float* save_mean, float* save_invstd, int64_t n,
                                     int64_t C, int training, float momentum, float eps, int relu,
                                     void* workspace, void* stream_) {
  return bn_forward_impl(x, residual, weight, bias, running_mean, running_var, y, save_mean, save_invstd, nullptr, nullptr,
                         n, C, training, momentum, eps, relu, workspace, stream_);
}

// As ddf_sparse_bn_forward; additionally writes y in the bf16x3 operand layout of the tcgen05 convs (split, same bytes
// as y; C % 32 == 0) and, when rounded != NULL, the tf32-rounded copy of y the wgrad kernels read.
extern "C" int ddf_sparse_bn_forward_split(const float* x, const float* residual, const float* weight,
                                           const float* bias, float* running_mean, float* running_var, float* y,
                                           float* save_mean, float* save_invstd, void* split, float* rounded,
                                           int64_t n, int64_t C, int training, float momentum, float eps, int relu,
                                           void* workspace, void* stream_) {
  DDF_CHECK_ARG(split != nullptr && C % 32 == 0, "sparse_bn_forward_split: needs the split buffer and C %% 32 == 0");
  DDF_CHECK_ARG(aligned16(split) && aligned16(rounded), "sparse_bn_forward_split: misaligned pointer");
  return bn_forward_impl(x, residual, weight, bias, running_mean, running_var, y, save_mean, save_invstd, split, rounded, n,
                         C, training, momentum, eps, relu, workspace, stream_);
}

static int bn_forward_impl(const float* x, const float* residual, const float* weight, const float* bias,
                           float* running_mean, float* running_var, float* y, float* save_mean, float* save_invstd,
                           void* split, float* rounded, int64_t n, int64_t C, int training, float momentum, float eps,
                           int relu, void* workspace, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(n >= 0 && pow2(C) && C >= 4 && C <= 1024,
                "sparse_bn_forward: C must be a power of two in [4, 1024], got n=%lld C=%lld",
                (long long)n, (long long)C);
  if (n == 0) return DDF_OK;
  DDF_CHECK_ARG(x && y, "sparse_bn_forward: null pointer");
  DDF_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(residual) && aligned16(weight) && aligned16(bias),
                "sparse_bn_forward: tensors must be 16-byte aligned");
  const long long n4 = n * C / 4;
  if (training) {
    DDF_CHECK_ARG(workspace && save_mean && save_invstd, "sparse_bn_forward: training needs workspace and save buffers");
    DDF_LAUNCH_PDL(bn_stats_kernel, stats_grid(n, C), kThreads, 2 * C * sizeof(double), stream, x, (int)n,
               (int)C, (double*)workspace, ws_part2(workspace, C), ws_counter(workspace, C), save_mean, save_invstd,
               running_mean, running_var, momentum, eps);
    if (split)
      DDF_LAUNCH_PDL(bn_apply_split_kernel, (unsigned)ddf::cdiv(n4 / 2, kThreads), kThreads, 0, stream, x, residual,
                 (const float*)save_mean, (const float*)save_invstd, weight, bias, y, reinterpret_cast<uint8_t*>(split),
                 rounded, n4 / 2, (int)C, relu, 0, eps);
    else
      DDF_LAUNCH_PDL(bn_apply_kernel, (unsigned)ddf::cdiv(n4, kThreads), kThreads, 0, stream, x, residual,
                 (const float*)save_mean, (const float*)save_invstd, weight, bias, y, n4, (int)C, relu, 0, eps);
  } else {
    DDF_CHECK_ARG(running_mean && running_var, "sparse_bn_forward: eval mode needs running statistics");
    if (split)
      DDF_LAUNCH_PDL(bn_apply_split_kernel, (unsigned)ddf::cdiv(n4 / 2, kThreads), kThreads, 0, stream, x, residual,
                 (const float*)running_mean, (const float*)running_var, weight, bias, y,
                 reinterpret_cast<uint8_t*>(split), rounded, n4 / 2, (int)C, relu, 1, eps);
    else
      DDF_LAUNCH_PDL(bn_apply_kernel, (unsigned)ddf::cdiv(n4, kThreads), kThreads, 0, stream, x, residual,
                 (const float*)running_mean, (const float*)running_var, weight, bias, y, n4, (int)C, relu, 1, eps);
  }
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// mean / invstd: the statistics forward normalised with (batch statistics in training; running mean
// and 1/sqrt(running_var + eps) in eval).  grad_x and grad_residual may be NULL when not needed.
extern "C" int ddf_sparse_bn_backward(const float* grad_y, const float* y, const float* x,
                                      const float* weight, const float* mean, const float* invstd,
                                      float* grad_x, float* grad_residual, float* grad_weight,
                                      float* grad_bias, int64_t n, int64_t C, int training, int relu,
                                      void* workspace, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(n >= 0 && pow2(C) && C >= 4 && C <= 1024,
                "sparse_bn_backward: C must be a power of two in [4, 1024], got n=%lld C=%lld",
                (long long)n, (long long)C);
  if (n == 0) {
    if (grad_weight) DDF_CUDA(cudaMemsetAsync(grad_weight, 0, sizeof(float) * C, stream));
    if (grad_bias) DDF_CUDA(cudaMemsetAsync(grad_bias, 0, sizeof(float) * C, stream));
    return DDF_OK;
  }
  DDF_CHECK_ARG(grad_y && x && mean && invstd && workspace && (y || !relu), "sparse_bn_backward: null pointer");
  DDF_CHECK_ARG(aligned16(grad_y) && aligned16(y) && aligned16(x) && aligned16(grad_x) && aligned16(grad_residual) &&
                    aligned16(weight) && aligned16(mean) && aligned16(invstd),
                "sparse_bn_backward: tensors must be 16-byte aligned");
  float* coef = ws_coef(workspace, C);
  DDF_LAUNCH_PDL(bn_bwd_reduce_kernel, stats_grid(n, C), kThreads, 2 * C * sizeof(double), stream, grad_y, y, x,
             mean, invstd, (int)n, (int)C, relu, training, (double*)workspace, ws_part2(workspace, C),
             ws_counter(workspace, C),
             grad_weight, grad_bias, coef);
  if (grad_x || grad_residual) {
    const long long n4 = n * C / 4;
    DDF_LAUNCH_PDL(bn_bwd_apply_kernel, (unsigned)ddf::cdiv(n4, kThreads), kThreads, 0, stream, grad_y, y, x, mean,
               invstd, weight, (const float*)coef, grad_x, grad_residual, n4, (int)C, relu);
  }
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}
