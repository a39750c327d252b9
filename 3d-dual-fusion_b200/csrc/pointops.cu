// Point-set ops of the 3D local self-attention (LocalTransformer): D-FPS, ball query, grouping,
// gathering — sm_100a.
//
// Reference kernels (<proj> = TransFusion/mmdet3d, CenterPoint/det3d, VoxelRCNN/pcdet/ops):
//   <proj>/ops/furthest_point_sample/src/furthest_point_sample_cuda.cu:25-141
//   <proj>/ops/ball_query/src/ball_query_cuda.cu:11-54
//   <proj>/ops/group_points/src/group_points_cuda.cu:10-31 (grad), :56-79 (fwd)
//   <proj>/ops/gather_points/src/gather_points_cuda.cu:8-26 (fwd), :51-70 (grad)
//
// FPS: the reference re-reads xyz and the running min-distance array from global memory in each of
// the npoint-1 dependent iterations and reduces with a 10-level __syncthreads tree.  Here a row's
// coordinates and min-distances live in registers for the whole kernel, the per-iteration arg-max
// is one packed 64-bit key (distance bits | inverted tie-break) reduced by warp shuffles and one
// shared-memory stage: 2 barriers per iteration instead of ~12 and no global traffic in the loop.
// The winner is bit-identical to the reference's: largest distance; among EQUAL distances (common on voxel-centre
// lattices) the reference's pairwise tree (stride B/2 ... 1, the lower slot keeps a tie) ends up preferring the
// thread whose index has the smallest BIT-REVERSED value, and inside a thread the strided scan keeps the lowest
// index: ties go to the smallest bitrev_log2(B)(index mod B), then the lowest index, B = the reference's block
// size (largest power of two <= n, capped at 1024).  Checked against the reference CUDA kernel itself
// (tests/test_reference_cuda_gpu.py).
// Squared distances use the contraction the reference kernels get from nvcc (default -fmad=true; read off the SASS
// of oracle/_ref/furthest_point_sample_ext.so and ball_query_ext.so built from the reference sources for sm_100):
// d = fma(dz, dz, fma(dx, dx, dy * dy)).  With it the sampled indices are identical to the reference CUDA kernel's
// on real clouds (tools/bench_reference_kernels.py, tests/test_reference_cuda_gpu.py); an uncontracted sum differs in
// the last bit often enough to pick another arg-max after a few hundred picks.
#include "common.cuh"

namespace {

constexpr int kFpsThreads = 1024;

__device__ __forceinline__ float sqdist(float x1, float y1, float z1, float x2, float y2, float z2) {
  const float dx = __fsub_rn(x2, x1), dy = __fsub_rn(y2, y1), dz = __fsub_rn(z2, z1);
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// bit reversal of (k mod B) in log2(B) bits (B a power of two); an involution, so it also decodes
__device__ __forceinline__ unsigned tie_major(unsigned k, int B) {
  if (B <= 1) return 0u;
  return __brev(k & (unsigned)(B - 1)) >> (__clz(B) + 1);
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int o) {
  unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
  lo = __shfl_xor_sync(0xffffffffu, lo, o);
  hi = __shfl_xor_sync(0xffffffffu, hi, o);
  return ((unsigned long long)hi << 32) | lo;
}

// One thread-block CLUSTER per row: CS CTAs x 1024 threads, PPT points per thread in registers
// (n <= PPT * CS * 1024).  Per iteration every CTA reduces its local arg-max, the owning thread
// pushes (key, x, y, z) into the exchange slot of every CTA of the cluster through distributed
// shared memory, one cluster barrier, and every CTA reads the CS candidates locally.
constexpr int kMaxCluster = 8;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(const void* local_smem, uint32_t rank) {
  uint32_t l = (uint32_t)__cvta_generic_to_shared(local_smem), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(l), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_u64(uint32_t addr, unsigned long long v) {
  asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

struct FpsSlot {
  unsigned long long key;
  float x, y, z, pad;
};

template <int PPT>
__global__ void __launch_bounds__(kFpsThreads)
fps_kernel(const float* __restrict__ xyz, float* __restrict__ temp, int* __restrict__ idxs, int n,
           int m, int ref_block, int cs) {
  ddf::pdl_sync();
  __shared__ unsigned long long s_key[kFpsThreads / 32];
  __shared__ FpsSlot s_slot[2][kMaxCluster];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_ctarank();
  const int row_id = blockIdx.x / cs;
  const float* row = xyz + (long long)row_id * n * 3;
  float* trow = temp ? temp + (long long)row_id * n : nullptr;
  int* out = idxs + (long long)row_id * m;

  float px[PPT], py[PPT], pz[PPT], pd[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int k = tid + kFpsThreads * (rank + cs * i);
    if (k < n) {
      px[i] = row[3 * k];
      py[i] = row[3 * k + 1];
      pz[i] = row[3 * k + 2];
      pd[i] = trow ? trow[k] : 1e10f;
    } else {
      px[i] = py[i] = pz[i] = 0.f;
      pd[i] = 0.f;
    }
  }
  float x1 = row[0], y1 = row[1], z1 = row[2];
  if (rank == 0 && tid == 0) out[0] = 0;
  cluster_barrier();  // every CTA of the cluster is resident before any remote store
  for (int j = 1; j < m; ++j) {
    const int buf = j & 1;
    unsigned long long best = 0ull;
    int best_i = 0;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const int k = tid + kFpsThreads * (rank + cs * i);
      if (k < n) {
        const float d = fminf(sqdist(x1, y1, z1, px[i], py[i], pz[i]), pd[i]);
        pd[i] = d;
        // smaller tie value wins: bitrev(k mod B) major, k / B minor; +1 keeps every real key non-zero
        const unsigned tie = 0x7fffffffu - ((tie_major(k, ref_block) << 21) | (unsigned)(k / ref_block));
        const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (tie + 1u);
        if (key > best) {
          best = key;
          best_i = i;
        }
      }
    }
    unsigned long long wbest = best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = shfl_xor_u64(wbest, o);
      wbest = other > wbest ? other : wbest;
    }
    if (lane == 0) s_key[warp] = wbest;
    __syncthreads();
    unsigned long long b2 = s_key[lane];  // kFpsThreads / 32 == 32 warps
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = shfl_xor_u64(b2, o);
      b2 = other > b2 ? other : b2;
    }
    // the owner of the CTA's best (keys are unique per point) publishes it to every CTA
    if ((b2 != 0ull && best == b2) || (b2 == 0ull && tid == 0)) {
      float wx = 0.f, wy = 0.f, wz = 0.f;
#pragma unroll
      for (int i = 0; i < PPT; ++i)
        if (i == best_i) {
          wx = px[i];
          wy = py[i];
          wz = pz[i];
        }
      for (int r = 0; r < cs; ++r) {
        const uint32_t a = map_to_cta(&s_slot[buf][rank], (uint32_t)r);
        st_cluster_u64(a, b2);
        st_cluster_f32(a + 8, wx);
        st_cluster_f32(a + 12, wy);
        st_cluster_f32(a + 16, wz);
      }
    }
    cluster_barrier();
    unsigned long long g = 0ull;
    int gi = 0;
    for (int r = 0; r < cs; ++r) {
      const unsigned long long kr = s_slot[buf][r].key;
      if (kr > g) {
        g = kr;
        gi = r;
      }
    }
    x1 = s_slot[buf][gi].x;
    y1 = s_slot[buf][gi].y;
    z1 = s_slot[buf][gi].z;
    if (rank == 0 && tid == 0) {
      const unsigned t = 0x7fffffffu - ((unsigned)(g & 0xffffffffu) - 1u);
      out[j] = (int)((t & 0x1fffffu) * (unsigned)ref_block + tie_major(t >> 21, ref_block));
    }
  }
  if (trow) {
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const int k = tid + kFpsThreads * (rank + cs * i);
      if (k < n) trow[k] = pd[i];
    }
  }
  cluster_barrier();  // no CTA exits while peers may still address its shared memory
}

// Rows of at most 8 * 1024 points (every 3D-DF configuration: <= 26000 queries over 6 cameras, 20000 for KITTI's
// single camera uses the cluster kernel): ONE CTA, points in registers, ONE __syncthreads per pick.  Every warp
// publishes its best (key, x, y, z) into a double-buffered slot before the barrier; after it every thread reduces the
// 32 warp keys with shuffles and reads the winner's coordinates from the winning warp's slot.  (The reference pays a
// 10-level __syncthreads tree and re-reads the cloud from global memory per pick.)
template <int PPT>
__global__ void __launch_bounds__(kFpsThreads)
fps_single_kernel(const float* __restrict__ xyz, float* __restrict__ temp, int* __restrict__ idxs, int n, int m,
                  int ref_block) {
  ddf::pdl_sync();
  __shared__ FpsSlot s_warp[2][kFpsThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = xyz + (long long)blockIdx.x * n * 3;
  float* trow = temp ? temp + (long long)blockIdx.x * n : nullptr;
  int* out = idxs + (long long)blockIdx.x * m;
  float px[PPT], py[PPT], pz[PPT], pd[PPT];
  unsigned tie[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int k = tid + kFpsThreads * i;
    if (k < n) {
      px[i] = row[3 * k];
      py[i] = row[3 * k + 1];
      pz[i] = row[3 * k + 2];
      pd[i] = trow ? trow[k] : 1e10f;
      // smaller tie value wins: bitrev(k mod B) major, k / B minor; +1 keeps every real key non-zero
      tie[i] = 0x7fffffffu - ((tie_major(k, ref_block) << 21) | (unsigned)(k / ref_block)) + 1u;
    } else {
      px[i] = py[i] = pz[i] = pd[i] = 0.f;
      tie[i] = 0u;
    }
  }
  float x1 = row[0], y1 = row[1], z1 = row[2];
  if (tid == 0) out[0] = 0;
  for (int j = 1; j < m; ++j) {
    const int buf = j & 1;
    unsigned long long best = 0ull;
    float bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      if (tie[i]) {
        const float d = fminf(sqdist(x1, y1, z1, px[i], py[i], pz[i]), pd[i]);
        pd[i] = d;
        const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | tie[i];
        if (key > best) {
          best = key;
          bx = px[i];
          by = py[i];
          bz = pz[i];
        }
      }
    }
    unsigned long long wbest = best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = shfl_xor_u64(wbest, o);
      wbest = other > wbest ? other : wbest;
    }
    // keys are unique per point: exactly one lane owns the warp's best (lane 0 when the warp has no point)
    if ((wbest != 0ull && best == wbest) || (wbest == 0ull && lane == 0)) {
      FpsSlot sl;
      sl.key = wbest;
      sl.x = bx;
      sl.y = by;
      sl.z = bz;
      sl.pad = 0.f;
      s_warp[buf][warp] = sl;
    }
    __syncthreads();
    const unsigned long long mine = s_warp[buf][lane].key;
    unsigned long long g = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = shfl_xor_u64(g, o);
      g = other > g ? other : g;
    }
    const int w = __ffs(__ballot_sync(0xffffffffu, mine == g)) - 1;
    x1 = s_warp[buf][w].x;
    y1 = s_warp[buf][w].y;
    z1 = s_warp[buf][w].z;
    if (tid == 0) {
      const unsigned t = 0x7fffffffu - ((unsigned)(g & 0xffffffffu) - 1u);
      out[j] = (int)((t & 0x1fffffu) * (unsigned)ref_block + tie_major(t >> 21, ref_block));
    }
  }
  if (trow) {
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const int k = tid + kFpsThreads * i;
      if (k < n) trow[k] = pd[i];
    }
  }
}

template <int PPT>
int launch_fps(const float* xyz, float* temp, int* idx, int64_t B, int n, int m, int ref_block, int cs,
               cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * cs));
  cfg.blockDim = dim3(kFpsThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DDF_CUDA(cudaLaunchKernelEx(&cfg, fps_kernel<PPT>, xyz, temp, idx, n, m, ref_block, cs));
  ddf::note_launches(1);
  return DDF_OK;
}

// ---- ball query: one warp per centre, points streamed through shared memory --------------------
constexpr int kBqWarps = 8;
constexpr int kBqTile = 1024;  // points per smem tile

__global__ void __launch_bounds__(kBqWarps * 32)
ball_query_kernel(const float* __restrict__ new_xyz, const float* __restrict__ xyz,
                  int* __restrict__ idx, int n, int m, float min_r2, float max_r2, int nsample) {
  ddf::pdl_sync();
  __shared__ float s_pts[kBqTile * 3];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * kBqWarps + warp;
  const bool live = c < m;
  const float* ctr = new_xyz + ((long long)b * m + (live ? c : 0)) * 3;
  const float cx = ctr[0], cy = ctr[1], cz = ctr[2];
  const float* row = xyz + (long long)b * n * 3;
  int* out = idx + ((long long)b * m + c) * nsample;
  int cnt = 0, first = 0;
  bool done = !live;
  for (int t0 = 0; t0 < n; t0 += kBqTile) {
    const int tn = min(kBqTile, n - t0);
    __syncthreads();
    for (int e = threadIdx.x; e < tn * 3; e += kBqWarps * 32) s_pts[e] = row[(long long)t0 * 3 + e];
    __syncthreads();
    if (__syncthreads_and(done)) break;  // every centre of the block is full (uniform)
    if (done) continue;
    for (int k0 = 0; k0 < tn && cnt < nsample; k0 += 32) {
      const int k = k0 + lane;
      bool hit = false;
      if (k < tn) {
        const float d2 = sqdist(s_pts[3 * k], s_pts[3 * k + 1], s_pts[3 * k + 2], cx, cy, cz);
        hit = (d2 == 0.f) || (d2 >= min_r2 && d2 < max_r2);
      }
      const unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (mask) {
        if (cnt == 0) first = t0 + k0 + __ffs(mask) - 1;
        const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
        if (hit && pos < nsample) out[pos] = t0 + k;
        cnt += __popc(mask);
      }
    }
    if (cnt >= nsample) done = true;
  }
  // unfilled slots repeat the first hit; no hit at all leaves the caller's zeros (ball_query.py:36)
  if (live && cnt > 0)
    for (int l = min(cnt, nsample) + lane; l < nsample; l += 32) out[l] = first;
}

// ---- grouping / gathering -----------------------------------------------------------------------
// out[b, c, e] = feat[b, c, idx[b, e]]   (e over npoint*nsample, or npoint for gather_points)
__global__ void __launch_bounds__(256)
index_rows_kernel(const float* __restrict__ feat, const int* __restrict__ idx, float* __restrict__ out,
                  int C, int N, long long E) {
  ddf::pdl_sync();
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (e >= E) return;
  const int j = idx[(long long)b * E + e];
  out[((long long)b * C + c) * E + e] = feat[((long long)b * C + c) * N + j];
}

// grad_feat[b, c, idx[b, e]] += grad_out[b, c, e]
__global__ void __launch_bounds__(256)
index_rows_grad_kernel(const float* __restrict__ gout, const int* __restrict__ idx,
                       float* __restrict__ gfeat, int C, int N, long long E) {
  ddf::pdl_sync();
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (e >= E) return;
  const int j = idx[(long long)b * E + e];
  atomicAdd(gfeat + ((long long)b * C + c) * N + j, gout[((long long)b * C + c) * E + e]);
}

// ---- LocalTransformer.scatter, "unique" / "replace" rule (<proj>/models/model_utils/pointformer.py:319-347) ----
// Every voxel that occurs in the ball-query index tensor takes the transformed feature of its FIRST occurrence in
// flattened (group, slot) order (what the reference's unique + flip + scatter_ yields on the CPU; its CUDA scatter_
// with duplicate indices is nondeterministic).  first[b, n] = min position of voxel n.
__global__ void __launch_bounds__(256) fill_int_kernel(int* __restrict__ p, int v, long long n) {
  ddf::pdl_sync();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void __launch_bounds__(256)
first_occurrence_kernel(const int* __restrict__ idx, int* __restrict__ first, int N, long long E) {
  ddf::pdl_sync();
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  const int b = blockIdx.y;
  if (e >= E) return;
  atomicMin(first + (long long)b * N + idx[(long long)b * E + e], (int)e);
}

// out[b, c, n] = feats[b, c, first[b, n]] where voxel n was hit (first < E); other columns keep their value
__global__ void __launch_bounds__(256)
scatter_first_kernel(const float* __restrict__ feats, const int* __restrict__ first, float* __restrict__ out,
                     int C, int N, long long E) {
  ddf::pdl_sync();
  const int n = blockIdx.x * 256 + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (n >= N) return;
  const int f = first[(long long)b * N + n];
  if (f < E) out[((long long)b * C + c) * N + n] = feats[((long long)b * C + c) * E + f];
}

// backward: grad_feats[b, c, first[b, n]] = grad_out[b, c, n] (positions are distinct: no atomics; grad_feats is
// zeroed by the caller), grad_features[b, c, n] = hit ? 0 : grad_out[b, c, n]
__global__ void __launch_bounds__(256)
scatter_first_grad_kernel(const float* __restrict__ gout, const int* __restrict__ first,
                          float* __restrict__ gfeats, float* __restrict__ gfeatures, int C, int N, long long E) {
  ddf::pdl_sync();
  const int n = blockIdx.x * 256 + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (n >= N) return;
  const int f = first[(long long)b * N + n];
  const long long o = ((long long)b * C + c) * N + n;
  const float g = gout[o];
  if (f < E) gfeats[((long long)b * C + c) * E + f] = g;
  gfeatures[o] = f < E ? 0.f : g;
}

int launch_index_rows(const float* feat, const int* idx, float* out, int64_t B, int64_t C, int64_t N,
                      int64_t E, cudaStream_t stream, bool grad) {
  if (B * C * E == 0) return DDF_OK;
  DDF_CHECK_ARG(C <= 65535 && B <= 65535, "group/gather: C and B must be <= 65535");
  dim3 grid((unsigned)ddf::cdiv(E, 256), (unsigned)C, (unsigned)B);
  if (grad)
    DDF_LAUNCH(index_rows_grad_kernel, grid, 256, 0, stream, feat, idx, out, (int)C, (int)N, (long long)E);
  else
    DDF_LAUNCH(index_rows_kernel, grid, 256, 0, stream, feat, idx, out, (int)C, (int)N, (long long)E);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

}  // namespace

// furthest_point_sampling_wrapper(B, N, m, xyz (B,N,3), temp (B,N) or NULL (=1e10), idx (B,m) out)
extern "C" int ddf_furthest_point_sampling(const float* xyz, float* temp, int* idx, int64_t B,
                                           int64_t N, int64_t m, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(B >= 0 && N >= 0 && m >= 0, "furthest_point_sampling: bad sizes");
  if (B == 0 || m == 0) return DDF_OK;
  DDF_CHECK_ARG(N > 0 && xyz && idx, "furthest_point_sampling: empty point set / null pointer");
  DDF_CHECK_ARG(N <= 8 * kMaxCluster * kFpsThreads, "furthest_point_sampling: at most %d points per row",
                8 * kMaxCluster * kFpsThreads);
  // the reference's block size decides its tie-break (opt_n_threads, furthest_point_sample_cuda.cu:9-13)
  int ref_block = 1;
  while (ref_block * 2 <= N && ref_block < 1024) ref_block *= 2;
  if (N <= 8 * kFpsThreads) {
    // one CTA per row, one barrier per pick
    const unsigned grid = (unsigned)B;
#define DDF_FPS_SINGLE(PPT) \
    DDF_LAUNCH(fps_single_kernel<PPT>, grid, kFpsThreads, 0, stream, xyz, temp, idx, (int)N, (int)m, ref_block)
    if (N <= kFpsThreads) DDF_FPS_SINGLE(1);
    else if (N <= 2 * kFpsThreads) DDF_FPS_SINGLE(2);
    else if (N <= 4 * kFpsThreads) DDF_FPS_SINGLE(4);
    else DDF_FPS_SINGLE(8);
#undef DDF_FPS_SINGLE
    DDF_LAUNCH_CHECK();
    return DDF_OK;
  }
  // cluster size: up to 4 points per thread first, then grow the per-thread count
  int cs = 1;
  while (cs < kMaxCluster && (long long)cs * 4 * kFpsThreads < N) cs *= 2;
  const long long per_thread = ddf::cdiv(N, (long long)cs * kFpsThreads);
  int rc;
  if (per_thread <= 1) rc = launch_fps<1>(xyz, temp, idx, B, (int)N, (int)m, ref_block, cs, stream);
  else if (per_thread <= 2) rc = launch_fps<2>(xyz, temp, idx, B, (int)N, (int)m, ref_block, cs, stream);
  else if (per_thread <= 4) rc = launch_fps<4>(xyz, temp, idx, B, (int)N, (int)m, ref_block, cs, stream);
  else rc = launch_fps<8>(xyz, temp, idx, B, (int)N, (int)m, ref_block, cs, stream);
  if (rc) return rc;
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// ball_query_wrapper(B, N, m, min_r, max_r, nsample, new_xyz (B,m,3), xyz (B,N,3), idx (B,m,nsample))
// idx must be zero-initialised by the caller exactly as in the reference (centres without any
// hit keep zeros).
extern "C" int ddf_ball_query(const float* new_xyz, const float* xyz, int* idx, int64_t B, int64_t N,
                              int64_t m, float min_radius, float max_radius, int64_t nsample,
                              void* stream_) {
  DDF_CHECK_ARG(B >= 0 && N >= 0 && m >= 0 && nsample > 0, "ball_query: bad sizes");
  DDF_CHECK_ARG(min_radius < max_radius, "ball_query: min_radius must be < max_radius");
  if (B * m == 0 || N == 0) return DDF_OK;
  DDF_CHECK_ARG(new_xyz && xyz && idx, "ball_query: null pointer");
  DDF_CHECK_ARG(B <= 65535, "ball_query: B must be <= 65535");
  dim3 grid((unsigned)ddf::cdiv(m, kBqWarps), (unsigned)B);
  DDF_LAUNCH(ball_query_kernel, grid, kBqWarps * 32, 0, (cudaStream_t)stream_, new_xyz, xyz, idx, (int)N,
             (int)m, min_radius * min_radius, max_radius * max_radius, (int)nsample);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// group_points forward(B, C, N, npoints, nsample, features (B,C,N), idx (B,npoints,nsample), out (B,C,npoints,nsample))
extern "C" int ddf_group_points(const float* features, const int* idx, float* out, int64_t B, int64_t C,
                                int64_t N, int64_t npoints, int64_t nsample, void* stream) {
  DDF_CHECK_ARG(B >= 0 && C >= 0 && N >= 0 && npoints >= 0 && nsample >= 0, "group_points: bad sizes");
  return launch_index_rows(features, idx, out, B, C, N, npoints * nsample, (cudaStream_t)stream, false);
}

// group_points backward: grad_features (B,C,N) zeroed inside, += grad_out (B,C,npoints,nsample)
extern "C" int ddf_group_points_grad(const float* grad_out, const int* idx, float* grad_features,
                                     int64_t B, int64_t C, int64_t N, int64_t npoints, int64_t nsample,
                                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(B >= 0 && C >= 0 && N >= 0 && npoints >= 0 && nsample >= 0, "group_points_grad: bad sizes");
  if (B * C * N > 0) {
    DDF_CHECK_ARG(grad_features != nullptr, "group_points_grad: null grad_features");
    DDF_CUDA(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)(B * C * N), stream));
  }
  return launch_index_rows(grad_out, idx, grad_features, B, C, N, npoints * nsample, stream, true);
}

// gather_points_wrapper(B, C, N, npoints, points (B,C,N), idx (B,npoints), out (B,C,npoints))
extern "C" int ddf_gather_points(const float* points, const int* idx, float* out, int64_t B, int64_t C,
                                 int64_t N, int64_t npoints, void* stream) {
  DDF_CHECK_ARG(B >= 0 && C >= 0 && N >= 0 && npoints >= 0, "gather_points: bad sizes");
  return launch_index_rows(points, idx, out, B, C, N, npoints, (cudaStream_t)stream, false);
}

extern "C" int ddf_gather_points_grad(const float* grad_out, const int* idx, float* grad_points,
                                      int64_t B, int64_t C, int64_t N, int64_t npoints, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(B >= 0 && C >= 0 && N >= 0 && npoints >= 0, "gather_points_grad: bad sizes");
  if (B * C * N > 0) {
    DDF_CHECK_ARG(grad_points != nullptr, "gather_points_grad: null grad_points");
    DDF_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)(B * C * N), stream));
  }
  return launch_index_rows(grad_out, idx, grad_points, B, C, N, npoints, stream, true);
}

// first [B, N] int32 <- first flattened position (< E = npoints * nsample) of every voxel in idx [B, E]; E = not hit
extern "C" int ddf_first_occurrence(const int* idx, int* first, int64_t B, int64_t N, int64_t E, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(B >= 0 && N >= 0 && E >= 0 && E < (1ll << 31) && B <= 65535, "first_occurrence: bad sizes");
  if (B * N == 0) return DDF_OK;
  DDF_CHECK_ARG(first != nullptr, "first_occurrence: null first");
  // fill with E: 0x7f7f7f7f would also do, but E keeps the "hit" test exact for any E
  DDF_LAUNCH(fill_int_kernel, (unsigned)ddf::cdiv(B * N, 256), 256, 0, stream, first, (int)E, (long long)(B * N));
  if (E == 0) return DDF_OK;
  DDF_CHECK_ARG(idx != nullptr, "first_occurrence: null idx");
  dim3 grid((unsigned)ddf::cdiv(E, 256), (unsigned)B);
  DDF_LAUNCH(first_occurrence_kernel, grid, 256, 0, stream, idx, first, (int)N, (long long)E);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// inout [B, C, N]: columns of hit voxels replaced by feats [B, C, E] at their first occurrence
extern "C" int ddf_scatter_first(const float* feats, const int* first, float* inout, int64_t B, int64_t C, int64_t N,
                                 int64_t E, void* stream_) {
  DDF_CHECK_ARG(B >= 0 && C >= 0 && N >= 0 && E >= 0 && C <= 65535 && B <= 65535, "scatter_first: bad sizes");
  if (B * C * N == 0 || E == 0) return DDF_OK;
  DDF_CHECK_ARG(feats && first && inout, "scatter_first: null pointer");
  dim3 grid((unsigned)ddf::cdiv(N, 256), (unsigned)C, (unsigned)B);
  DDF_LAUNCH(scatter_first_kernel, grid, 256, 0, (cudaStream_t)stream_, feats, first, inout, (int)C, (int)N, (long long)E);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

extern "C" int ddf_scatter_first_grad(const float* grad_out, const int* first, float* grad_feats,
                                      float* grad_features, int64_t B, int64_t C, int64_t N, int64_t E,
                                      void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DDF_CHECK_ARG(B >= 0 && C >= 0 && N >= 0 && E >= 0 && C <= 65535 && B <= 65535, "scatter_first_grad: bad sizes");
  if (B * C * E > 0) {
    DDF_CHECK_ARG(grad_feats != nullptr, "scatter_first_grad: null grad_feats");
    DDF_CUDA(cudaMemsetAsync(grad_feats, 0, sizeof(float) * (size_t)(B * C * E), stream));
  }
  if (B * C * N == 0) return DDF_OK;
  DDF_CHECK_ARG(grad_out && first && grad_features, "scatter_first_grad: null pointer");
  dim3 grid((unsigned)ddf::cdiv(N, 256), (unsigned)C, (unsigned)B);
  DDF_LAUNCH(scatter_first_grad_kernel, grid, 256, 0, stream, grad_out, first, grad_feats, grad_features, (int)C, (int)N,
             (long long)E);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}
