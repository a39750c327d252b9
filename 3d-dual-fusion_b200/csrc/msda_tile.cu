// Dual-query deformable attention, tile-staged (the hot-path form of MSDA: one feature level, 4 points).
//
// Replaces, in ONE kernel per direction,
//   softmax over the L*P logits + sampling_locations = ref + offsets / (W, H)
//       <proj>/models/model_utils/ops/modules/ms_deform_attn.py:149-166
//   MSDeformAttnFunction forward / backward
//       <proj>/models/model_utils/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299, 301-403
// for the shapes every shipped 3D-DF config uses (L = 1, P = 4, D = 8 or 16, fp32); any other shape goes
// through the generic kernels in msda.cu behind the reference's own op signature.
//
// Design (north_star: "TMA-staged image tiles in shared memory, vectorised loads, warp-shuffle reductions"):
//   * plan (once per encoder forward, reference points are shared by all layers and by backward): queries are
//     binned by the 16 x 16-pixel tile of the feature map their reference point falls in (counting sort on the
//     device: histogram with atomic ranks -> single-block scan -> scatter) and cut into work items of at most
//     128 queries of one tile.
//   * a CTA owns (work item, head group).  One thread issues ONE 4-D TMA box load
//     (cp.async.bulk.tensor.4d, UTMALDG): the tile plus a 6-pixel halo, 28 x 28 pixels x 128 bytes (= 2 heads of
//     16 channels or 4 heads of 8) = 98 KB, two CTAs per SM so one CTA's load overlaps the other's sampling.
//     Out-of-image pixels are zero-filled by TMA, which IS the bilinear zero-padding rule: corners need no
//     bounds predicates at all.
//   * sampling: 8 lanes per query (head-in-group x 4-channel slice), 4 queries per warp.  Each lane computes the
//     corner geometry of ONE sampling point (D = 16) and the softmax of its logit (shuffle max / sum over the
//     4 lanes of the head); the 4 points are then broadcast by shuffles (pixel index + 4 weights) and every
//     lane reads its 16-byte slice of the 4 corners with LDS.128 at compile-time offsets from one base.
//     A quarter warp (2 heads x 4 lanes, or 4 x 2) covers all 32 banks once: conflict-free.
//   * a sampling point that leaves the staged window (offsets beyond the halo) falls back to bounds-checked
//     global loads for that point only.
//   * backward: same staging for the value reads; grad_value by red.global.add.v4.f32 (measured: shared-memory
//     atomics are slower per lane than REDG on this part), grad wrt the RAW offsets / logits (softmax backward
//     in registers), channel reductions by shuffles.
// Arithmetic that decides which pixels are touched replicates the module's fp32 sequence exactly
// (off / W, + ref, * W, - 0.5, floor; no FMA contraction).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kTile = 16;
constexpr int kHalo = 6;
constexpr int kStage = kTile + 2 * kHalo;          // 28 staged pixels per side
constexpr int kPixBytes = 128;                     // bytes per staged pixel = one head group
constexpr int kStageBytes = kStage * kStage * kPixBytes;
constexpr int kQC = 128;                           // queries per work item
constexpr int kThreads = 256;
constexpr int kP = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------------
// plan: int32 buffer
//   [0] n_work | counts[NT] | tile_start[NT + 1] | q_tile[NQ] | q_rank[NQ] | perm[NQ] | work[3 * NWmax]
// ------------------------------------------------------------------------------------------------
struct PlanLayout {
  long long counts, tile_start, q_tile, q_rank, perm, work, total;
  int NT, NWmax, TX, TY;
};
__host__ __device__ inline PlanLayout plan_layout(long long N, long long NQ, int H, int W) {
  PlanLayout p;
  p.TX = (W + kTile - 1) / kTile;
  p.TY = (H + kTile - 1) / kTile;
  p.NT = (int)(N * p.TX * p.TY);
  p.NWmax = (int)(p.NT + NQ / kQC + 1);
  p.counts = 1;
  p.tile_start = p.counts + p.NT;
  p.q_tile = p.tile_start + p.NT + 1;
  p.q_rank = p.q_tile + NQ;
  p.perm = p.q_rank + NQ;
  p.work = p.perm + NQ;
  p.total = p.work + 3ll * p.NWmax;
  return p;
}

__global__ void __launch_bounds__(kThreads)
msda_plan_count_kernel(const float* __restrict__ ref, const int* __restrict__ qbatch, int* __restrict__ plan,
                       PlanLayout pl, long long NQ, int Lq, int H, int W) {
  ddf::pdl_sync();
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= NQ) return;
  const float2 r = __ldg(reinterpret_cast<const float2*>(ref) + i);
  // tile of the reference pixel; reference points outside [0, 1) (padded rows, unseen voxels) are clamped
  int px = (int)floorf(r.x * (float)W), py = (int)floorf(r.y * (float)H);
  px = min(max(px, 0), W - 1);
  py = min(max(py, 0), H - 1);
  const int b = qbatch ? qbatch[i] : (int)(i / Lq);     // ragged query list: the image of every query is given
  const int t = (b * pl.TY + py / kTile) * pl.TX + px / kTile;
  plan[pl.q_tile + i] = t;
  plan[pl.q_rank + i] = atomicAdd(plan + pl.counts + t, 1);
}

// single block: tile_start = exclusive scan of counts; work items = (tile, first query slot, count <= kQC)
__global__ void __launch_bounds__(1024) msda_plan_scan_kernel(int* __restrict__ plan, PlanLayout pl) {
  ddf::pdl_sync();
  __shared__ int s_q[32], s_w[32];
  __shared__ int carry_q, carry_w;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_q = carry_w = 0;
  __syncthreads();
  for (int base = 0; base < pl.NT; base += 1024) {
    const int t = base + threadIdx.x;
    const int c = t < pl.NT ? plan[pl.counts + t] : 0;
    const int nw = (c + kQC - 1) / kQC;
    int iq = c, iw = nw;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, iq, o), b2 = __shfl_up_sync(0xffffffffu, iw, o);
      if (lane >= o) { iq += a; iw += b2; }
    }
    if (lane == 31) { s_q[warp] = iq; s_w[warp] = iw; }
    __syncthreads();
    if (warp == 0) {
      int a = s_q[lane], b2 = s_w[lane];
      int ia = a, ib = b2;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, ia, o), y = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += x; ib += y; }
      }
      s_q[lane] = ia - a;
      s_w[lane] = ib - b2;
    }
    __syncthreads();
    const int q0 = carry_q + s_q[warp] + iq - c;      // exclusive
    const int w0 = carry_w + s_w[warp] + iw - nw;
    if (t < pl.NT) {
      plan[pl.tile_start + t] = q0;
      for (int j = 0; j < nw; ++j) {
        int* wk = plan + pl.work + 3ll * (w0 + j);
        wk[0] = t;
        wk[1] = q0 + j * kQC;
        wk[2] = min(kQC, c - j * kQC);
      }
    }
    __syncthreads();
    if (threadIdx.x == 1023) { carry_q = q0 + c; carry_w = w0 + nw; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    plan[pl.tile_start + pl.NT] = carry_q;
    plan[0] = carry_w;
  }
}

__global__ void __launch_bounds__(kThreads)
msda_plan_scatter_kernel(int* __restrict__ plan, PlanLayout pl, long long NQ) {
  ddf::pdl_sync();
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= NQ) return;
  plan[pl.perm + plan[pl.tile_start + plan[pl.q_tile + i]] + plan[pl.q_rank + i]] = (int)i;
}

// ------------------------------------------------------------------------------------------------
// shared pieces of the forward / backward kernels
// ------------------------------------------------------------------------------------------------
struct PointGeom {
  int code;       // >= 0: pixel index of corner (h_low, w_low) inside the staged window; -1: contributes nothing;
                  // -2: outside the staged window -> global fallback
  int hw;         // (h_low << 16) | (w_low & 0xffff)
  float lh, lw, hh, hw_;
};

__device__ __forceinline__ PointGeom point_geom(float ref_x, float ref_y, float off_x, float off_y, int H, int W,
                                                int x0, int y0) {
  PointGeom g;
  // module arithmetic: loc = ref + off / (W, H) (ms_deform_attn.py:155-160), then the kernel's
  // h_im = loc_y * H - 0.5 (ms_deform_im2col_cuda.cuh:285-286); no FMA contraction anywhere
  const float loc_x = __fadd_rn(ref_x, __fdiv_rn(off_x, (float)W));
  const float loc_y = __fadd_rn(ref_y, __fdiv_rn(off_y, (float)H));
  const float h_im = __fsub_rn(__fmul_rn(loc_y, (float)H), 0.5f);
  const float w_im = __fsub_rn(__fmul_rn(loc_x, (float)W), 0.5f);
  const bool inr = (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);
  const float hf = floorf(h_im), wf = floorf(w_im);
  const int h_low = inr ? (int)hf : 0, w_low = inr ? (int)wf : 0;
  g.lh = h_im - hf;
  g.lw = w_im - wf;
  g.hh = 1.f - g.lh;
  g.hw_ = 1.f - g.lw;
  g.hw = (h_low << 16) | (w_low & 0xffff);
  const int ry = h_low - y0, rx = w_low - x0;
  const bool staged = ry >= 0 && rx >= 0 && ry + 1 < kStage && rx + 1 < kStage;
  g.code = !inr ? -1 : (staged ? ry * kStage + rx : -2);
  return g;
}

__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4fma(float a, float4 v, float4 acc) {
  acc.x = fmaf(a, v.x, acc.x);
  acc.y = fmaf(a, v.y, acc.y);
  acc.z = fmaf(a, v.z, acc.z);
  acc.w = fmaf(a, v.w, acc.w);
  return acc;
}
__device__ __forceinline__ float f4dot(float4 a, float4 b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// the four corners of one point for this lane's 4 channels: from the staged window, or (code == -2) from global
// memory with per-corner bounds checks
__device__ __forceinline__ void load_corners(int code, int hw, const uint8_t* stage_lane, const float* vglob, int H,
                                             int W, long long pix, float4 (&v)[4]) {
  if (code >= 0) {
    const float4* base = reinterpret_cast<const float4*>(stage_lane + code * kPixBytes);
    v[0] = base[0];
    v[1] = base[kPixBytes / 16];
    v[2] = base[kStage * kPixBytes / 16];
    v[3] = base[(kStage + 1) * kPixBytes / 16];
  } else if (code == -2) {
    const int h_low = hw >> 16, w_low = (int)(short)(hw & 0xffff);
    const bool hl = h_low >= 0, hh = h_low + 1 <= H - 1, wl = w_low >= 0, wh = w_low + 1 <= W - 1;
    const float* p = vglob + ((long long)h_low * W + w_low) * pix;
    v[0] = (hl && wl) ? ldg4(p) : f4zero();
    v[1] = (hl && wh) ? ldg4(p + pix) : f4zero();
    v[2] = (hh && wl) ? ldg4(p + (long long)W * pix) : f4zero();
    v[3] = (hh && wh) ? ldg4(p + (long long)(W + 1) * pix) : f4zero();
  } else {
    v[0] = v[1] = v[2] = v[3] = f4zero();
  }
}

__device__ __forceinline__ void stage_tile(uint8_t* stage, uint64_t* bar, const CUtensorMap* map, int c0, int x0,
                                           int y0, int b) {
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(kStageBytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(stage)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(x0), "r"(y0), "r"(b)
        : "memory");
  }
}
// The same window staged by LDGSTS instead of TMA (kept selectable for A/B measurements): every thread copies the
// 16-byte chunk (tid & 7) of the pixels (tid >> 3) + 32 i; out-of-image pixels are zero-filled (src-size 0).
__device__ __forceinline__ void stage_tile_cp_async(uint8_t* stage, const float* __restrict__ value, int b, int H,
                                                    int W, long long pix, int c0, int x0, int y0) {
  const int part = threadIdx.x & 7;
  int yy = (threadIdx.x >> 3) / kStage, xx = (threadIdx.x >> 3) % kStage;
  const float* img = value + (long long)b * H * W * pix + c0 + part * 4;
  uint32_t dst = smem_u32(stage) + (uint32_t)((threadIdx.x >> 3) * kPixBytes + part * 16);
#pragma unroll 5
  for (int i = 0; i < (kStage * kStage + 31) / 32; ++i) {
    if (yy < kStage) {
      const int gy = y0 + yy, gx = x0 + xx;
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const float* src = img + (ok ? ((long long)gy * W + gx) * pix : 0);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u)
                   : "memory");
    }
    dst += 32 * kPixBytes;
    xx += 32 - kStage;          // 32 pixels further along the row-major window
    yy += 1;
    if (xx >= kStage) {
      xx -= kStage;
      yy += 1;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void wait_tile_cp_async() {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
}

__device__ __forceinline__ void wait_tile(uint64_t* bar) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done, spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr)
        : "memory");
    if (!done && ++spins > (1u << 24)) {
      printf("msda tile: TMA never completed (block %d,%d)\n", blockIdx.x, blockIdx.y);
      __trap();
    }
  } while (!done);
}

// ------------------------------------------------------------------------------------------------
// forward.  TPH lanes per head (D = 4 * TPH channels), HPC = 8 / TPH heads per CTA, 4 / TPH points per lane.
// ------------------------------------------------------------------------------------------------
template <int TPH, bool TMA>
__global__ void __launch_bounds__(kThreads, 2)
msda_tile_fwd_kernel(const __grid_constant__ CUtensorMap map_value, const float* __restrict__ value,
                     const float* __restrict__ ref, const float* __restrict__ off, const float* __restrict__ logit,
                     const int* __restrict__ plan, long long perm_off, long long work_off, float* __restrict__ out,
                     int H, int W, int M, int Lq, int TX, int TY) {
  ddf::pdl_sync();
  constexpr int D = 4 * TPH, HPC = 8 / TPH, PPL = kP / TPH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ uint64_t bar;
  if ((int)blockIdx.x >= plan[0]) return;
  const int* wk = plan + work_off + 3ll * blockIdx.x;
  const int tile = wk[0], q_begin = wk[1], q_count = wk[2];
  const int b = tile / (TX * TY), ty = (tile / TX) % TY, tx = tile % TX;
  const int x0 = tx * kTile - kHalo, y0 = ty * kTile - kHalo;
  const int hg = blockIdx.y;                       // head group
  if constexpr (TMA)
    stage_tile(stage, &bar, &map_value, hg * (kPixBytes / 4), x0, y0, b);
  else
    stage_tile_cp_async(stage, value, b, H, W, (long long)M * (4 * TPH), hg * (kPixBytes / 4), x0, y0);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & 7, qi = lane >> 3;
  const int hl = sub / TPH, c4 = sub % TPH;
  const int m = hg * HPC + hl;
  const int grp = lane & ~(TPH - 1);               // first lane of this head's lane group
  const long long pix = (long long)M * D;
  const float* vglob = value + (long long)b * H * W * pix + m * D + c4 * 4;
  const uint8_t* stage_lane = stage + hl * (D * 4) + c4 * 16;
  const int* perm = plan + perm_off + q_begin;
  if constexpr (TMA) __syncthreads();   // barrier initialised before anyone polls it

  // All per-query inputs of the CTA's (at most kQC / 32 = 4) rounds are requested up front: two dependent global
  // latencies (query index -> reference point / offsets / logits) per CTA instead of three per round, and they
  // overlap the TMA load of the tile.
  constexpr int R = kQC / 32;
  int gq[R];
  float2 rf[R], o[R][PPL];
  float lg[R][PPL];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int slot = r * 32 + warp * 4 + qi;
    gq[r] = perm[slot < q_count ? slot : q_count - 1];          // global query index b * Lq + q
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (r * 32 >= q_count) break;
    rf[r] = __ldg(reinterpret_cast<const float2*>(ref) + gq[r]);
    const long long hbase = ((long long)gq[r] * M + m) * kP;
#pragma unroll
    for (int s = 0; s < PPL; ++s) {
      o[r][s] = __ldg(reinterpret_cast<const float2*>(off) + hbase + c4 * PPL + s);
      lg[r][s] = __ldg(logit + hbase + c4 * PPL + s);
    }
  }
  if constexpr (TMA) wait_tile(&bar); else wait_tile_cp_async();
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (r * 32 >= q_count) break;
    const bool live = r * 32 + warp * 4 + qi < q_count;
    // this lane's points: geometry + softmax
    PointGeom g[PPL];
    float pr[PPL];
#pragma unroll
    for (int s = 0; s < PPL; ++s) g[s] = point_geom(rf[r].x, rf[r].y, o[r][s].x, o[r][s].y, H, W, x0, y0);
    float mx = lg[r][0];
#pragma unroll
    for (int s = 1; s < PPL; ++s) mx = fmaxf(mx, lg[r][s]);
#pragma unroll
    for (int sh = 1; sh < TPH; sh <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sh));
    float sum = 0.f;
#pragma unroll
    for (int s = 0; s < PPL; ++s) {
      pr[s] = expf(lg[r][s] - mx);
      sum += pr[s];
    }
#pragma unroll
    for (int sh = 1; sh < TPH; sh <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
    float wgt[PPL][4];
#pragma unroll
    for (int s = 0; s < PPL; ++s) {
      const float a = pr[s] / sum;
      wgt[s][0] = a * (g[s].hh * g[s].hw_);
      wgt[s][1] = a * (g[s].hh * g[s].lw);
      wgt[s][2] = a * (g[s].lh * g[s].hw_);
      wgt[s][3] = a * (g[s].lh * g[s].lw);
    }
    float4 acc = f4zero();
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      const int src = grp + p / PPL, s = p % PPL;
      const int code = __shfl_sync(0xffffffffu, g[s].code, src);
      const int hw = __shfl_sync(0xffffffffu, g[s].hw, src);
      const float w1 = __shfl_sync(0xffffffffu, wgt[s][0], src), w2 = __shfl_sync(0xffffffffu, wgt[s][1], src);
      const float w3 = __shfl_sync(0xffffffffu, wgt[s][2], src), w4 = __shfl_sync(0xffffffffu, wgt[s][3], src);
      float4 v[4];
      load_corners(code, hw, stage_lane, vglob, H, W, pix, v);
      acc = f4fma(w1, v[0], acc);
      acc = f4fma(w2, v[1], acc);
      acc = f4fma(w3, v[2], acc);
      acc = f4fma(w4, v[3], acc);
    }
    if (live) *reinterpret_cast<float4*>(out + ((long long)gq[r] * M + m) * D + c4 * 4) = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// backward: grad_value (red.global), grad wrt raw offsets [.., P, 2] and logits [.., P]
// ------------------------------------------------------------------------------------------------
template <int TPH, bool TMA>
__global__ void __launch_bounds__(kThreads, 2)
msda_tile_bwd_kernel(const __grid_constant__ CUtensorMap map_value, const float* __restrict__ value,
                     const float* __restrict__ ref, const float* __restrict__ off, const float* __restrict__ logit,
                     const float* __restrict__ gout, const int* __restrict__ plan, long long perm_off,
                     long long work_off, float* __restrict__ gvalue, float* __restrict__ goff,
                     float* __restrict__ glogit, int H, int W, int M, int Lq, int TX, int TY) {
  ddf::pdl_sync();
  constexpr int D = 4 * TPH, HPC = 8 / TPH, PPL = kP / TPH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ uint64_t bar;
  if ((int)blockIdx.x >= plan[0]) return;
  const int* wk = plan + work_off + 3ll * blockIdx.x;
  const int tile = wk[0], q_begin = wk[1], q_count = wk[2];
  const int b = tile / (TX * TY), ty = (tile / TX) % TY, tx = tile % TX;
  const int x0 = tx * kTile - kHalo, y0 = ty * kTile - kHalo;
  const int hg = blockIdx.y;
  if constexpr (TMA)
    stage_tile(stage, &bar, &map_value, hg * (kPixBytes / 4), x0, y0, b);
  else
    stage_tile_cp_async(stage, value, b, H, W, (long long)M * (4 * TPH), hg * (kPixBytes / 4), x0, y0);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & 7, qi = lane >> 3;
  const int hl = sub / TPH, c4 = sub % TPH;
  const int m = hg * HPC + hl;
  const int grp = lane & ~(TPH - 1);
  const long long pix = (long long)M * D;
  const long long img = (long long)b * H * W * pix + m * D + c4 * 4;
  const float* vglob = value + img;
  float* gvglob = gvalue + img;
  const uint8_t* stage_lane = stage + hl * (D * 4) + c4 * 16;
  const int* perm = plan + perm_off + q_begin;
  if constexpr (TMA) __syncthreads();

  constexpr int R = kQC / 32;
  int gq[R];
  float2 rf[R], o[R][PPL];
  float lg[R][PPL];
  float4 gor[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int slot = r * 32 + warp * 4 + qi;
    gq[r] = perm[slot < q_count ? slot : q_count - 1];
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (r * 32 >= q_count) break;
    rf[r] = __ldg(reinterpret_cast<const float2*>(ref) + gq[r]);
    const long long hbase = ((long long)gq[r] * M + m) * kP;
#pragma unroll
    for (int s = 0; s < PPL; ++s) {
      o[r][s] = __ldg(reinterpret_cast<const float2*>(off) + hbase + c4 * PPL + s);
      lg[r][s] = __ldg(logit + hbase + c4 * PPL + s);
    }
    gor[r] = ldg4(gout + ((long long)gq[r] * M + m) * D + c4 * 4);
  }
  if constexpr (TMA) wait_tile(&bar); else wait_tile_cp_async();
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (r * 32 >= q_count) break;
    const bool live = r * 32 + warp * 4 + qi < q_count;
    const long long hbase = ((long long)gq[r] * M + m) * kP;
    const float4 go = live ? gor[r] : f4zero();
    PointGeom g[PPL];
    float pr[PPL];
#pragma unroll
    for (int s = 0; s < PPL; ++s) g[s] = point_geom(rf[r].x, rf[r].y, o[r][s].x, o[r][s].y, H, W, x0, y0);
    float mx = lg[r][0];
#pragma unroll
    for (int s = 1; s < PPL; ++s) mx = fmaxf(mx, lg[r][s]);
#pragma unroll
    for (int sh = 1; sh < TPH; sh <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sh));
    float sum = 0.f;
#pragma unroll
    for (int s = 0; s < PPL; ++s) {
      pr[s] = expf(lg[r][s] - mx);
      sum += pr[s];
    }
#pragma unroll
    for (int sh = 1; sh < TPH; sh <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
#pragma unroll
    for (int s = 0; s < PPL; ++s) pr[s] = pr[s] / sum;
    float ga_own[PPL], gw_own[PPL], gh_own[PPL];
#pragma unroll
    for (int s = 0; s < PPL; ++s) ga_own[s] = gw_own[s] = gh_own[s] = 0.f;
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      const int src = grp + p / PPL, s = p % PPL;
      const int code = __shfl_sync(0xffffffffu, g[s].code, src);
      const int hw = __shfl_sync(0xffffffffu, g[s].hw, src);
      const float lh = __shfl_sync(0xffffffffu, g[s].lh, src), lw = __shfl_sync(0xffffffffu, g[s].lw, src);
      const float a = __shfl_sync(0xffffffffu, pr[s], src);
      const float hh = 1.f - lh, hw_ = 1.f - lw;
      float4 v[4];
      load_corners(code, hw, stage_lane, vglob, H, W, pix, v);
      const float w1 = hh * hw_, w2 = hh * lw, w3 = lh * hw_, w4 = lh * lw;
      const float4 tg = make_float4(go.x * a, go.y * a, go.z * a, go.w * a);
      if (live && code != -1) {
        const int h_low = hw >> 16, w_low = (int)(short)(hw & 0xffff);
        const bool hl_ = h_low >= 0, hh_ = h_low + 1 <= H - 1, wl = w_low >= 0, wh = w_low + 1 <= W - 1;
        float* gp = gvglob + ((long long)h_low * W + w_low) * pix;
        if (hl_ && wl) red_add_v4(gp, w1 * tg.x, w1 * tg.y, w1 * tg.z, w1 * tg.w);
        if (hl_ && wh) red_add_v4(gp + pix, w2 * tg.x, w2 * tg.y, w2 * tg.z, w2 * tg.w);
        if (hh_ && wl) red_add_v4(gp + (long long)W * pix, w3 * tg.x, w3 * tg.y, w3 * tg.z, w3 * tg.w);
        if (hh_ && wh) red_add_v4(gp + (long long)(W + 1) * pix, w4 * tg.x, w4 * tg.y, w4 * tg.z, w4 * tg.w);
      }
      // val, d/dw, d/dh per channel (ms_deform_im2col_cuda.cuh:116-158)
      float4 val = f4zero(), dw = f4zero(), dh = f4zero();
      val = f4fma(w1, v[0], val);
      val = f4fma(w2, v[1], val);
      val = f4fma(w3, v[2], val);
      val = f4fma(w4, v[3], val);
      dw = f4fma(-hh, v[0], dw);
      dw = f4fma(hh, v[1], dw);
      dw = f4fma(-lh, v[2], dw);
      dw = f4fma(lh, v[3], dw);
      dh = f4fma(-hw_, v[0], dh);
      dh = f4fma(-lw, v[1], dh);
      dh = f4fma(hw_, v[2], dh);
      dh = f4fma(lw, v[3], dh);
      float ga = f4dot(go, val);       // d out / d attention weight
      float gw = f4dot(tg, dw);        // d out / d (pixel x) = d / d offset_x  (loc_x * W = ref_x * W + off_x)
      float gh = f4dot(tg, dh);
#pragma unroll
      for (int o = 1; o < TPH; o <<= 1) {
        ga += __shfl_xor_sync(0xffffffffu, ga, o);
        gw += __shfl_xor_sync(0xffffffffu, gw, o);
        gh += __shfl_xor_sync(0xffffffffu, gh, o);
      }
      if (lane == src) {
        ga_own[s] = ga;
        gw_own[s] = gw;
        gh_own[s] = gh;
      }
    }
    // softmax backward over the P logits of the head: gl_p = a_p * (ga_p - sum_j a_j ga_j)
    float dot = 0.f;
#pragma unroll
    for (int s = 0; s < PPL; ++s) dot = fmaf(pr[s], ga_own[s], dot);
#pragma unroll
    for (int o = 1; o < TPH; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (live) {
#pragma unroll
      for (int s = 0; s < PPL; ++s) {
        const int p = c4 * PPL + s;
        glogit[hbase + p] = pr[s] * (ga_own[s] - dot);
        reinterpret_cast<float2*>(goff)[hbase + p] = make_float2(gw_own[s], gh_own[s]);
      }
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------
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

// value [N, H, W, M * D] fp32 as a 4-D tensor (c, x, y, n); box = 32 floats x 28 x 28 x 1, zero fill outside
bool make_value_map(CUtensorMap* m, const float* value, int64_t N, int H, int W, int C) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {kPixBytes / 4, kStage, kStage, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(value), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// staging of the value window: one TMA box load (default) or LDGSTS from all threads (DDF_MSDA_STAGE=cp; also the
// fallback when the driver has no cuTensorMapEncodeTiled).  Measured on B200, C-TF shape: TMA 0.099 ms, LDGSTS
// 0.126 ms per forward launch (profiles/r2_msda.md).
bool stage_with_tma() {
  static const bool tma = [] {
    const char* e = getenv("DDF_MSDA_STAGE");
    return !(e && e[0] == 'c') && encode_fn() != nullptr;
  }();
  return tma;
}

int check_shapes(int64_t N, int64_t H, int64_t W, int64_t M, int64_t D, int64_t NQ, const char* what) {
  DDF_CHECK_ARG(N >= 0 && H > 0 && W > 0 && M > 0 && NQ >= 0, "%s: bad sizes", what);
  DDF_CHECK_ARG((D == 8 || D == 16) && (M * D) % 32 == 0,
                "%s: the tile-staged kernel takes D in {8, 16} with M * D a multiple of 32 (D=%lld M=%lld)", what,
                (long long)D, (long long)M);
  DDF_CHECK_ARG(H < 32768 && W < 32768 && N * H * W * M * D < (1ll << 40) && NQ < (1ll << 31),
                "%s: problem too large", what);
  return DDF_OK;
}

}  // namespace

// Supported by the tile-staged kernels? (one level, 4 points, fp32, D in {8, 16}, M * D % 32 == 0)
extern "C" int ddf_msda_tile_supported(int64_t M, int64_t D, int64_t L, int64_t P) {
  return L == 1 && P == kP && (D == 8 || D == 16) && (M * D) % 32 == 0;
}

extern "C" int64_t ddf_msda_plan_bytes(int64_t N, int64_t NQ, int64_t H, int64_t W) {
  if (N < 0 || NQ < 0 || H <= 0 || W <= 0) return -1;
  return plan_layout(N, NQ, (int)H, (int)W).total * 4;
}

// reference_points [NQ, 2] (x, y in [0, 1]) -> plan (ddf_msda_plan_bytes bytes, int32 aligned).  Regular layout:
// query_batch = NULL and query i belongs to image i / Lq (NQ = N * Lq).  Ragged layout (only the real queries of a
// zero-padded per-camera layout): query_batch [NQ] int32 gives the image of every query, Lq is ignored.
extern "C" int ddf_msda_plan(const float* reference_points, const int* query_batch, void* plan_, int64_t N, int64_t NQ,
                             int64_t Lq, int64_t H, int64_t W, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(N >= 0 && NQ >= 0 && H > 0 && W > 0 && NQ < (1ll << 31), "msda_plan: bad sizes");
  DDF_CHECK_ARG(query_batch || (Lq > 0 && NQ == N * Lq) || NQ == 0, "msda_plan: NQ must be N * Lq without query_batch");
  DDF_CHECK_ARG(plan_ != nullptr, "msda_plan: null plan");
  int* plan = reinterpret_cast<int*>(plan_);
  const PlanLayout pl = plan_layout(N, NQ, (int)H, (int)W);
  DDF_CUDA(cudaMemsetAsync(plan, 0, sizeof(int) * (size_t)(pl.tile_start), stream));   // n_work + counts
  if (NQ == 0) return DDF_OK;
  DDF_CHECK_ARG(reference_points != nullptr, "msda_plan: null reference_points");
  const unsigned grid = (unsigned)ddf::cdiv(NQ, kThreads);
  DDF_LAUNCH(msda_plan_count_kernel, grid, kThreads, 0, stream, reference_points, query_batch, plan, pl, NQ,
             (int)(Lq > 0 ? Lq : 1), (int)H, (int)W);
  DDF_LAUNCH(msda_plan_scan_kernel, 1, 1024, 0, stream, plan, pl);
  DDF_LAUNCH(msda_plan_scatter_kernel, grid, kThreads, 0, stream, plan, pl, NQ);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// out [NQ, M*D] = MSDA(value [N, H*W, M, D]; loc = ref + offsets / (W, H); weights = softmax(logits))
// offsets [NQ, M, 1, 4, 2] raw (pixels), logits [NQ, M, 4]; NQ and the query -> image map are those of the plan.
extern "C" int ddf_msda_tile_forward(const float* value, const float* reference_points, const float* offsets,
                                     const float* logits, const void* plan_, float* out, int64_t N, int64_t H,
                                     int64_t W, int64_t M, int64_t D, int64_t NQ, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_shapes(N, H, W, M, D, NQ, "msda_tile_forward");
  if (rc) return rc;
  if (N * NQ == 0) return DDF_OK;
  DDF_CHECK_ARG(value && reference_points && offsets && logits && plan_ && out, "msda_tile_forward: null pointer");
  CUtensorMap map;
  const bool tma = stage_with_tma();
  if (tma && !make_value_map(&map, value, N, (int)H, (int)W, (int)(M * D))) {
    ddf::set_error("msda_tile_forward: cuTensorMapEncodeTiled failed");
    return DDF_ERR_CUDA;
  }
  const PlanLayout pl = plan_layout(N, NQ, (int)H, (int)W);
  const int* plan = reinterpret_cast<const int*>(plan_);
  const int smem = kStageBytes + 128;
  const dim3 grid((unsigned)pl.NWmax, (unsigned)(M * D / 32));
#define DDF_TILE_FWD(TPH, TMA)                                                                                   \
  do {                                                                                                            \
    DDF_SET_SMEM_ONCE((msda_tile_fwd_kernel<TPH, TMA>), smem);                                                    \
    DDF_LAUNCH_PDL((msda_tile_fwd_kernel<TPH, TMA>), grid, kThreads, smem, stream, map, value, reference_points,      \
               offsets, logits, plan, pl.perm, pl.work, out, (int)H, (int)W, (int)M, (int)NQ, pl.TX, pl.TY);      \
  } while (0)
  if (D == 16) {
    if (tma) DDF_TILE_FWD(4, true); else DDF_TILE_FWD(4, false);
  } else {
    if (tma) DDF_TILE_FWD(2, true); else DDF_TILE_FWD(2, false);
  }
#undef DDF_TILE_FWD
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// grad_value [N, H*W, M, D] (zeroed inside), grad_offsets [N, Lq, M, 1, 4, 2], grad_logits [N, Lq, M, 4]
extern "C" int ddf_msda_tile_backward(const float* value, const float* reference_points, const float* offsets,
                                      const float* logits, const float* grad_out, const void* plan_,
                                      float* grad_value, float* grad_offsets, float* grad_logits, int64_t N,
                                      int64_t H, int64_t W, int64_t M, int64_t D, int64_t NQ, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_shapes(N, H, W, M, D, NQ, "msda_tile_backward");
  if (rc) return rc;
  if (N > 0) {
    DDF_CHECK_ARG(grad_value != nullptr, "msda_tile_backward: null grad_value");
    DDF_CUDA(cudaMemsetAsync(grad_value, 0, sizeof(float) * (size_t)(N * H * W * M * D), stream));
  }
  if (N * NQ == 0) return DDF_OK;
  DDF_CHECK_ARG(value && reference_points && offsets && logits && grad_out && plan_ && grad_offsets && grad_logits,
                "msda_tile_backward: null pointer");
  CUtensorMap map;
  const bool tma = stage_with_tma();
  if (tma && !make_value_map(&map, value, N, (int)H, (int)W, (int)(M * D))) {
    ddf::set_error("msda_tile_backward: cuTensorMapEncodeTiled failed");
    return DDF_ERR_CUDA;
  }
  const PlanLayout pl = plan_layout(N, NQ, (int)H, (int)W);
  const int* plan = reinterpret_cast<const int*>(plan_);
  const int smem = kStageBytes + 128;
  const dim3 grid((unsigned)pl.NWmax, (unsigned)(M * D / 32));
#define DDF_TILE_BWD(TPH, TMA)                                                                                   \
  do {                                                                                                            \
    DDF_SET_SMEM_ONCE((msda_tile_bwd_kernel<TPH, TMA>), smem);                                                    \
    DDF_LAUNCH_PDL((msda_tile_bwd_kernel<TPH, TMA>), grid, kThreads, smem, stream, map, value, reference_points,      \
               offsets, logits, grad_out, plan, pl.perm, pl.work, grad_value, grad_offsets, grad_logits, (int)H,  \
               (int)W, (int)M, (int)NQ, pl.TX, pl.TY);                                                            \
  } while (0)
  if (D == 16) {
    if (tma) DDF_TILE_BWD(4, true); else DDF_TILE_BWD(4, false);
  } else {
    if (tma) DDF_TILE_BWD(2, true); else DDF_TILE_BWD(2, false);
  }
#undef DDF_TILE_BWD
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}
