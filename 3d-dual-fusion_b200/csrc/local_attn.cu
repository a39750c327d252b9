// 3D local self-attention core of the LocalTransformer: softmax(Q K^T / sqrt(hd)) V inside every group of 32 grouped
// voxel tokens, per head — forward and backward — on token-major [T, 3C] projections.
//
// Replaces what nn.MultiheadAttention does between its in_proj and out_proj for the reference's
//   TransformerEncoderLayerPreNorm   <proj>/models/model_utils/pointformer.py:10-44
//   LocalTransformer.forward          <proj>/models/model_utils/pointformer.py:349-380
// where the sequences are the B' * npoint ball-query groups of nsample = 32 tokens (1.57 M tokens per layer for the
// CenterPoint config).  The reference path permutes the (B', C, npoint, 32) tensor to (32, B' * npoint, C) and back,
// and the library attention materialises per-head [B' * npoint * heads, 32, 32] score tensors plus several
// transposed copies of Q / K / V (0.8 GB each): measured 43 ms per training step on B200 for ONE layer.
//
// Here the whole LocalTransformer stays token-major ([T, C] rows = (group, slot)), the projections are plain row-major
// GEMMs, and this kernel is the only place that looks inside a group:
//   one warp per (group, head); lane i owns query / output row i; K and V (32 x hd) sit in shared memory and are
//   read as 16-byte broadcasts; scores, the softmax and P V never leave registers.
//   backward recomputes P, lane i produces dQ_i; P and dS go through a padded [32][33] shared tile so that lane j can
//   then sum the columns for dK_j and dV_j (no atomics, deterministic).
// HBM-bound: forward reads 3C and writes C floats per token, backward reads 4C and writes 3C.
#include "common.cuh"

namespace {

constexpr int kNS = 32;            // tokens per group = lanes per warp
constexpr int kWarps = 4;          // (group, head) pairs per CTA

template <int HD>
__global__ void __launch_bounds__(kWarps * 32)
local_attn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, long long n_pairs, int heads, float scale) {
  ddf::pdl_sync();
  __shared__ __align__(16) float s_k[kWarps][kNS][HD];
  __shared__ __align__(16) float s_v[kWarps][kNS][HD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pair = (long long)blockIdx.x * kWarps + warp;
  if (pair >= n_pairs) return;
  const long long group = pair / heads;
  const int h = (int)(pair % heads);
  const int C = heads * HD;
  const float* row = qkv + (group * kNS + lane) * 3ll * C + h * HD;
  float q[HD];
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    const float4 a = ldg4(row + d), b = ldg4(row + C + d), c = ldg4(row + 2 * C + d);
    q[d] = a.x * scale; q[d + 1] = a.y * scale; q[d + 2] = a.z * scale; q[d + 3] = a.w * scale;
    *reinterpret_cast<float4*>(&s_k[warp][lane][d]) = b;
    *reinterpret_cast<float4*>(&s_v[warp][lane][d]) = c;
  }
  __syncwarp();
  float s[kNS];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kNS; ++j) {
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 k4 = *reinterpret_cast<const float4*>(&s_k[warp][j][d]);
      acc = fmaf(q[d], k4.x, acc);
      acc = fmaf(q[d + 1], k4.y, acc);
      acc = fmaf(q[d + 2], k4.z, acc);
      acc = fmaf(q[d + 3], k4.w, acc);
    }
    s[j] = acc;
    mx = fmaxf(mx, acc);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kNS; ++j) {
    s[j] = expf(s[j] - mx);
    sum += s[j];
  }
  const float inv = 1.f / sum;
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < kNS; ++j) {
    const float p = s[j] * inv;
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 v4 = *reinterpret_cast<const float4*>(&s_v[warp][j][d]);
      o[d] = fmaf(p, v4.x, o[d]);
      o[d + 1] = fmaf(p, v4.y, o[d + 1]);
      o[d + 2] = fmaf(p, v4.z, o[d + 2]);
      o[d + 3] = fmaf(p, v4.w, o[d + 3]);
    }
  }
  float* dst = out + (group * kNS + lane) * (long long)C + h * HD;
#pragma unroll
  for (int d = 0; d < HD; d += 4) *reinterpret_cast<float4*>(dst + d) = make_float4(o[d], o[d + 1], o[d + 2], o[d + 3]);
}

template <int HD>
__global__ void __launch_bounds__(kWarps * 32)
local_attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ gout, float* __restrict__ gqkv,
                      long long n_pairs, int heads, float scale) {
  ddf::pdl_sync();
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per warp: Q (scaled), K, V, dO: 4 x [32][HD]; P, dS: 2 x [32][33]
  float* base = smem + warp * (4 * kNS * HD + 2 * kNS * 33);
  float (*s_q)[HD] = reinterpret_cast<float (*)[HD]>(base);
  float (*s_k)[HD] = reinterpret_cast<float (*)[HD]>(base + kNS * HD);
  float (*s_v)[HD] = reinterpret_cast<float (*)[HD]>(base + 2 * kNS * HD);
  float (*s_g)[HD] = reinterpret_cast<float (*)[HD]>(base + 3 * kNS * HD);
  float (*s_p)[33] = reinterpret_cast<float (*)[33]>(base + 4 * kNS * HD);
  float (*s_ds)[33] = reinterpret_cast<float (*)[33]>(base + 4 * kNS * HD + kNS * 33);
  const long long pair = (long long)blockIdx.x * kWarps + warp;
  if (pair >= n_pairs) return;
  const long long group = pair / heads;
  const int h = (int)(pair % heads);
  const int C = heads * HD;
  const long long t = group * kNS + lane;
  const float* row = qkv + t * 3ll * C + h * HD;
  const float* grow = gout + t * (long long)C + h * HD;
  float q[HD], go[HD];
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    const float4 a = ldg4(row + d), b = ldg4(row + C + d), c = ldg4(row + 2 * C + d), g = ldg4(grow + d);
    q[d] = a.x * scale; q[d + 1] = a.y * scale; q[d + 2] = a.z * scale; q[d + 3] = a.w * scale;
    go[d] = g.x; go[d + 1] = g.y; go[d + 2] = g.z; go[d + 3] = g.w;
    *reinterpret_cast<float4*>(&s_q[lane][d]) = make_float4(q[d], q[d + 1], q[d + 2], q[d + 3]);
    *reinterpret_cast<float4*>(&s_k[lane][d]) = b;
    *reinterpret_cast<float4*>(&s_v[lane][d]) = c;
    *reinterpret_cast<float4*>(&s_g[lane][d]) = g;
  }
  __syncwarp();
  // row i = lane: P_ij, dP_ij = dO_i . V_j, dS_ij = P_ij (dP_ij - sum_j P_ij dP_ij).  The row's scores / dP live in
  // the padded [32][33] tiles (conflict-free for row AND column access), not in registers.
  float mx = -INFINITY;
#pragma unroll 4
  for (int j = 0; j < kNS; ++j) {
    float acc = 0.f, acc2 = 0.f;
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 k4 = *reinterpret_cast<const float4*>(&s_k[j][d]);
      const float4 v4 = *reinterpret_cast<const float4*>(&s_v[j][d]);
      acc = fmaf(q[d], k4.x, acc); acc = fmaf(q[d + 1], k4.y, acc); acc = fmaf(q[d + 2], k4.z, acc); acc = fmaf(q[d + 3], k4.w, acc);
      acc2 = fmaf(go[d], v4.x, acc2); acc2 = fmaf(go[d + 1], v4.y, acc2); acc2 = fmaf(go[d + 2], v4.z, acc2); acc2 = fmaf(go[d + 3], v4.w, acc2);
    }
    s_p[lane][j] = acc;
    s_ds[lane][j] = acc2;
    mx = fmaxf(mx, acc);
  }
  float sum = 0.f;
#pragma unroll 8
  for (int j = 0; j < kNS; ++j) {
    const float e = expf(s_p[lane][j] - mx);
    s_p[lane][j] = e;
    sum += e;
  }
  const float inv = 1.f / sum;
  float dot = 0.f;
#pragma unroll 8
  for (int j = 0; j < kNS; ++j) {
    const float pj = s_p[lane][j] * inv;
    s_p[lane][j] = pj;
    dot = fmaf(pj, s_ds[lane][j], dot);
  }
  float dq[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) dq[d] = 0.f;
#pragma unroll 4
  for (int j = 0; j < kNS; ++j) {
    const float ds = s_p[lane][j] * (s_ds[lane][j] - dot);
    s_ds[lane][j] = ds;
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 k4 = *reinterpret_cast<const float4*>(&s_k[j][d]);
      dq[d] = fmaf(ds, k4.x, dq[d]); dq[d + 1] = fmaf(ds, k4.y, dq[d + 1]);
      dq[d + 2] = fmaf(ds, k4.z, dq[d + 2]); dq[d + 3] = fmaf(ds, k4.w, dq[d + 3]);
    }
  }
  float* gr = gqkv + t * 3ll * C + h * HD;
#pragma unroll
  for (int d = 0; d < HD; d += 4)   // d score / d q_unscaled = scale * d score / d q_scaled
    *reinterpret_cast<float4*>(gr + d) = make_float4(dq[d] * scale, dq[d + 1] * scale, dq[d + 2] * scale, dq[d + 3] * scale);
  __syncwarp();
  // column j = lane: dK_j = sum_i dS_ij q_i (q already carries the scale), dV_j = sum_i P_ij dO_i
  float dk[HD], dv[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) dk[d] = dv[d] = 0.f;
#pragma unroll 4
  for (int i = 0; i < kNS; ++i) {
    const float ds = s_ds[i][lane], pp = s_p[i][lane];
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 q4 = *reinterpret_cast<const float4*>(&s_q[i][d]);
      const float4 g4 = *reinterpret_cast<const float4*>(&s_g[i][d]);
      dk[d] = fmaf(ds, q4.x, dk[d]); dk[d + 1] = fmaf(ds, q4.y, dk[d + 1]);
      dk[d + 2] = fmaf(ds, q4.z, dk[d + 2]); dk[d + 3] = fmaf(ds, q4.w, dk[d + 3]);
      dv[d] = fmaf(pp, g4.x, dv[d]); dv[d + 1] = fmaf(pp, g4.y, dv[d + 1]);
      dv[d + 2] = fmaf(pp, g4.z, dv[d + 2]); dv[d + 3] = fmaf(pp, g4.w, dv[d + 3]);
    }
  }
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    *reinterpret_cast<float4*>(gr + C + d) = make_float4(dk[d], dk[d + 1], dk[d + 2], dk[d + 3]);
    *reinterpret_cast<float4*>(gr + 2 * C + d) = make_float4(dv[d], dv[d + 1], dv[d + 2], dv[d + 3]);
  }
}

template <int HD>
int launch_fwd(const float* qkv, float* out, long long n_pairs, int heads, cudaStream_t stream) {
  const float scale = 1.0f / sqrtf((float)HD);
  DDF_LAUNCH(local_attn_fwd_kernel<HD>, (unsigned)ddf::cdiv(n_pairs, kWarps), kWarps * 32, 0, stream, qkv, out, n_pairs,
             heads, scale);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

template <int HD>
int launch_bwd(const float* qkv, const float* gout, float* gqkv, long long n_pairs, int heads, cudaStream_t stream) {
  const float scale = 1.0f / sqrtf((float)HD);
  const int smem = kWarps * (4 * kNS * HD + 2 * kNS * 33) * 4;
  DDF_SET_SMEM_ONCE(local_attn_bwd_kernel<HD>, smem);
  DDF_LAUNCH(local_attn_bwd_kernel<HD>, (unsigned)ddf::cdiv(n_pairs, kWarps), kWarps * 32, smem, stream, qkv, gout, gqkv,
             n_pairs, heads, scale);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

int check(int64_t groups, int64_t heads, int64_t head_dim, int64_t group_size, const char* what) {
  DDF_CHECK_ARG(groups >= 0 && heads > 0 && groups * heads < (1ll << 40), "%s: bad sizes", what);
  DDF_CHECK_ARG(group_size == kNS && (head_dim == 16 || head_dim == 32),
                "%s: groups of 32 tokens with head_dim 16 or 32 (got group %lld, head_dim %lld)", what,
                (long long)group_size, (long long)head_dim);
  return DDF_OK;
}

}  // namespace

extern "C" int ddf_local_attn_supported(int64_t heads, int64_t head_dim, int64_t group_size) {
  return heads > 0 && group_size == kNS && (head_dim == 16 || head_dim == 32);
}

// qkv [groups * 32, 3 * heads * head_dim] (q | k | v, as nn.MultiheadAttention's in_proj emits them) ->
// out [groups * 32, heads * head_dim] = softmax(q k^T / sqrt(head_dim)) v inside every group, per head
extern "C" int ddf_local_attn_forward(const float* qkv, float* out, int64_t groups, int64_t heads, int64_t head_dim,
                                      int64_t group_size, void* stream_) {
  int rc = check(groups, heads, head_dim, group_size, "local_attn_forward");
  if (rc || groups == 0) return rc;
  DDF_CHECK_ARG(qkv && out, "local_attn_forward: null pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  return head_dim == 32 ? launch_fwd<32>(qkv, out, groups * heads, (int)heads, stream)
                        : launch_fwd<16>(qkv, out, groups * heads, (int)heads, stream);
}

// grad_qkv [groups * 32, 3 * C] from grad_out [groups * 32, C]; P is recomputed
extern "C" int ddf_local_attn_backward(const float* qkv, const float* grad_out, float* grad_qkv, int64_t groups,
                                       int64_t heads, int64_t head_dim, int64_t group_size, void* stream_) {
  int rc = check(groups, heads, head_dim, group_size, "local_attn_backward");
  if (rc || groups == 0) return rc;
  DDF_CHECK_ARG(qkv && grad_out && grad_qkv, "local_attn_backward: null pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  return head_dim == 32 ? launch_bwd<32>(qkv, grad_out, grad_qkv, groups * heads, (int)heads, stream)
                        : launch_bwd<16>(qkv, grad_out, grad_qkv, groups * heads, (int)heads, stream);
}
