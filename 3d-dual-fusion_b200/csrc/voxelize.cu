// Hard / dynamic voxelization for sm_100a.
//
// Contract (bit-exact with the reference, CPU and GPU paths agree there):
//   <TF>/ops/voxel/src/voxelization_cpu.cpp:44-142, voxelization_cuda.cu:25-61,106-180,184-326
//   - coordinate of a point: c = floor((p - range_min) / voxel_size) per axis, fp32 IEEE sub + div,
//     stored as (z, y, x); out-of-range points are dropped;
//   - voxels are numbered in FIRST-SEEN order of the input points;
//   - inside a voxel, points keep input order, at most max_points are stored;
//   - when a NEW voxel would be number max_voxels, processing stops there: that point and every
//     later point (even of already-open voxels) is dropped (cpu.cpp:73 / cuda.cu:164 `break`).
//
// The reference finds duplicates with an O(N^2) scan and numbers voxels on ONE thread
// (voxelization_cuda.cu:106-180) with four device synchronisations.  Here:
//   K1  point -> key, open-addressing hash insert (int32 CAS), and a concurrent SORTED insertion of
//       the point index into the voxel's max_points-slot list (atomicMin cascade: slot s ends up
//       holding the (s+1)-th smallest point index of the voxel whatever the thread interleaving);
//   K2  flag "point is the first of its voxel", exclusive scan -> first-seen voxel rank;
//   K3  locate the cut-off point (rank == max_voxels), publish voxel_num;
//   K4  one thread per (voxel, slot, feature): gather the point rows into voxels[M, T, F], write
//       coors and num_points.  Rows [0, voxel_num) are written completely (unused slots = 0), so the
//       caller does not need to pre-zero the max_voxels-sized buffers.
// No host synchronisation; voxel_num stays on the device (the Python shim reads it once).
#include <limits.h>

#include "common.cuh"
#include "scan.cuh"

namespace {

constexpr int kThreads = 256;

struct Grid3 {
  float vx, vy, vz, x0, y0, z0;
  int gx, gy, gz;
};

__device__ __forceinline__ unsigned hash32(unsigned k) {
  k ^= k >> 16;
  k *= 0x7feb352dU;
  k ^= k >> 15;
  k *= 0x846ca68bU;
  k ^= k >> 16;
  return k;
}

// returns linear key (z*gy + y)*gx + x, or -1 when the point is outside the grid
__device__ __forceinline__ int point_key(const float* p, const Grid3& g, int* cz, int* cy, int* cx) {
  const float fx = floorf(__fdiv_rn(__fsub_rn(p[0], g.x0), g.vx));
  const float fy = floorf(__fdiv_rn(__fsub_rn(p[1], g.y0), g.vy));
  const float fz = floorf(__fdiv_rn(__fsub_rn(p[2], g.z0), g.vz));
  // written so NaN coordinates fail the test
  if (!(fx >= 0.f && fx < (float)g.gx && fy >= 0.f && fy < (float)g.gy && fz >= 0.f &&
        fz < (float)g.gz))
    return -1;
  *cx = (int)fx;
  *cy = (int)fy;
  *cz = (int)fz;
  return (*cz * g.gy + *cy) * g.gx + *cx;
}

__global__ void __launch_bounds__(kThreads)
dynamic_voxelize_kernel(const float* __restrict__ points, int* __restrict__ coors, Grid3 g, int n,
                        int F) {
  ddf::pdl_sync();
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  int cz, cy, cx;
  const int key = point_key(points + (long long)i * F, g, &cz, &cy, &cx);
  int* o = coors + 3ll * i;
  if (key < 0) {
    o[0] = o[1] = o[2] = -1;  // CPU semantics (voxelization_cpu.cpp:33-38)
  } else {
    o[0] = cz;
    o[1] = cy;
    o[2] = cx;
  }
}

// K1
__global__ void __launch_bounds__(kThreads)
vox_insert_kernel(const float* __restrict__ points, Grid3 g, int n, int F, int T, int* keys,
                  unsigned mask, int* lists, int* __restrict__ slot_of_point) {
  ddf::pdl_sync();
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  int cz, cy, cx;
  const int key = point_key(points + (long long)i * F, g, &cz, &cy, &cx);
  if (key < 0) {
    slot_of_point[i] = -1;
    return;
  }
  unsigned h = hash32((unsigned)key) & mask;
  while (true) {
    const int k = atomicCAS(keys + h, -1, key);
    if (k == -1 || k == key) break;
    h = (h + 1) & mask;
  }
  slot_of_point[i] = (int)h;
  // sorted insertion: every value visits slots 0,1,... until it settles or falls off the end
  int* lst = lists + (long long)h * T;
  int v = i;
  for (int s = 0; s < T; ++s) {
    const int old = atomicMin(lst + s, v);
    if (old == 0x7f7f7f7f) break;  // slot was empty: v settled, nothing displaced
    if (old > v) v = old;          // v settled here, carry the displaced larger index on
  }
}

// K2
__global__ void __launch_bounds__(kThreads)
vox_flag_first_kernel(const int* __restrict__ slot_of_point, const int* __restrict__ lists, int T,
                      int n, int* __restrict__ is_first) {
  ddf::pdl_sync();
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const int h = slot_of_point[i];
  is_first[i] = (h >= 0 && lists[(long long)h * T] == i) ? 1 : 0;
}

// K3: first_point[rank] = i; cut = index of the first point whose voxel would be number max_voxels
__global__ void __launch_bounds__(kThreads)
vox_rank_kernel(const int* __restrict__ is_first, const int* __restrict__ rank, int n,
                int max_voxels, int* __restrict__ first_point, int* __restrict__ cut,
                int* __restrict__ voxel_num) {
  ddf::pdl_sync();
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i == 0) {
    const int total = rank[n];
    *voxel_num = total < max_voxels ? total : max_voxels;
    if (total <= max_voxels) *cut = n;  // no cut-off: every real index (< n) passes, the empty-slot sentinel does not
  }
  if (i >= n || !is_first[i]) return;
  const int r = rank[i];
  if (r < max_voxels) first_point[r] = i;
  if (r == max_voxels) *cut = i;
}

// K4
__global__ void __launch_bounds__(kThreads)
vox_gather_kernel(const float* __restrict__ points, Grid3 g, int F, int T,
                  const int* __restrict__ first_point, const int* __restrict__ slot_of_point,
                  const int* __restrict__ keys, const int* __restrict__ lists,
                  const int* __restrict__ cut_p, const int* __restrict__ voxel_num_p,
                  float* __restrict__ voxels, int* __restrict__ coors,
                  int* __restrict__ num_points, long long total) {
  ddf::pdl_sync();
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (t >= total) return;
  const int TF = T * F;
  const int v = (int)(t / TF);
  if (v >= *voxel_num_p) return;
  const int e = (int)(t % TF);
  const int s = e / F, f = e % F;
  const int h = slot_of_point[first_point[v]];
  const int cut = *cut_p;
  const int pi = lists[(long long)h * T + s];
  // lists are sorted ascending, so "pi < cut" also bounds the count
  voxels[t] = pi < cut ? points[(long long)pi * F + f] : 0.f;
  if (e == 0) {
    int cnt = 0;
    for (int k = 0; k < T; ++k) cnt += lists[(long long)h * T + k] < cut ? 1 : 0;
    num_points[v] = cnt;
    const int key = keys[h];
    const int x = key % g.gx, y = (key / g.gx) % g.gy, z = key / (g.gx * g.gy);
    coors[3 * v] = z;
    coors[3 * v + 1] = y;
    coors[3 * v + 2] = x;
  }
}

// K4', fused HardSimpleVFE (TransFusion/mmdet3d/models/voxel_encoders/voxel_encoder.py:27-44): instead of writing the
// padded [M, T, F] voxel tensor only for a reduction kernel to read it back, sum the (at most T) points of a voxel in
// list order and divide by the count.  One thread per (voxel, feature < NF); coors / num_points as in K4.
__global__ void __launch_bounds__(kThreads)
vox_mean_kernel(const float* __restrict__ points, Grid3 g, int F, int T, int NF,
                const int* __restrict__ first_point, const int* __restrict__ slot_of_point,
                const int* __restrict__ keys, const int* __restrict__ lists,
                const int* __restrict__ cut_p, const int* __restrict__ voxel_num_p,
                float* __restrict__ mean, int* __restrict__ coors, int* __restrict__ num_points, long long total) {
  ddf::pdl_sync();
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (t >= total) return;
  const int v = (int)(t / NF);
  if (v >= *voxel_num_p) return;
  const int f = (int)(t % NF);
  const int h = slot_of_point[first_point[v]];
  const int cut = *cut_p;
  float sum = 0.f;
  int cnt = 0;
  for (int k = 0; k < T; ++k) {
    const int pi = lists[(long long)h * T + k];
    if (pi < cut) {     // lists are sorted ascending: the valid entries come first
      sum += points[(long long)pi * F + f];
      ++cnt;
    }
  }
  mean[t] = sum / (float)cnt;
  if (f == 0) {
    num_points[v] = cnt;
    const int key = keys[h];
    const int x = key % g.gx, y = (key / g.gx) % g.gy, z = key / (g.gx * g.gy);
    coors[3 * v] = z;
    coors[3 * v + 1] = y;
    coors[3 * v + 2] = x;
  }
}

int make_grid(const float* voxel_size, const float* range, Grid3* g) {
  g->vx = voxel_size[0];
  g->vy = voxel_size[1];
  g->vz = voxel_size[2];
  g->x0 = range[0];
  g->y0 = range[1];
  g->z0 = range[2];
  DDF_CHECK_ARG(g->vx > 0 && g->vy > 0 && g->vz > 0, "voxelize: voxel_size must be positive");
  // reference: grid = round((max - min) / voxel) in float (voxelization_cuda.cu:206-208)
  g->gx = (int)roundf((range[3] - range[0]) / g->vx);
  g->gy = (int)roundf((range[4] - range[1]) / g->vy);
  g->gz = (int)roundf((range[5] - range[2]) / g->vz);
  DDF_CHECK_ARG(g->gx > 0 && g->gy > 0 && g->gz > 0, "voxelize: empty grid");
  DDF_CHECK_ARG((long long)g->gx * g->gy * g->gz < INT_MAX, "voxelize: grid exceeds int32 keys");
  return DDF_OK;
}

unsigned table_slots(long long n) {
  unsigned s = 1024;
  while ((long long)s < 2 * n) s <<= 1;
  return s;
}

struct VoxWs {
  int *keys, *lists, *slot, *is_first, *rank, *first_point, *scan_ws, *cut;
  size_t bytes;
};

VoxWs carve(void* base, long long n, int T, int max_voxels) {
  VoxWs w;
  const unsigned slots = table_slots(n);
  size_t off = 0;
  auto take = [&](size_t cnt) {
    int* p = base ? reinterpret_cast<int*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += ((cnt * sizeof(int) + 255) / 256) * 256;
    return p;
  };
  w.keys = take(slots);
  w.lists = take((size_t)slots * T);
  w.slot = take(n);
  w.is_first = take(n);
  w.rank = take(n + 1);
  w.first_point = take(max_voxels > 0 ? max_voxels : 1);
  w.scan_ws = take(ddf::scan_workspace_bytes(n) / sizeof(int));
  w.cut = take(1);
  w.bytes = off;
  return w;
}

}  // namespace

extern "C" int64_t ddf_hard_voxelize_workspace_bytes(int64_t num_points, int64_t max_points,
                                                     int64_t max_voxels) {
  if (num_points < 0 || max_points <= 0 || max_voxels < 0) return -1;
  return (int64_t)carve(nullptr, num_points, (int)max_points, (int)max_voxels).bytes;
}

extern "C" int ddf_dynamic_voxelize(const float* points, int* coors, const float* voxel_size_host,
                                    const float* coors_range_host, int64_t num_points,
                                    int64_t num_features, void* stream_) {
  Grid3 g;
  DDF_CHECK_ARG(voxel_size_host && coors_range_host, "dynamic_voxelize: null voxel_size/range");
  int rc = make_grid(voxel_size_host, coors_range_host, &g);
  if (rc) return rc;
  DDF_CHECK_ARG(num_points >= 0 && num_features >= 3, "dynamic_voxelize: need (N, >=3) points");
  if (num_points == 0) return DDF_OK;
  DDF_CHECK_ARG(points && coors, "dynamic_voxelize: null pointer");
  DDF_LAUNCH(dynamic_voxelize_kernel, (unsigned)ddf::cdiv(num_points, kThreads), kThreads, 0, (cudaStream_t)stream_, points, coors, g, (int)num_points,
                                                     (int)num_features);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

namespace {
// voxels != NULL: the padded voxel tensor (reference contract); else mean != NULL: per-voxel mean of the first
// mean_features point features (fused HardSimpleVFE)
int hard_voxelize_impl(const float* points, float* voxels, float* mean, int64_t mean_features, int* coors,
                       int* num_points_per_voxel, int* voxel_num,
                       const float* voxel_size_host, const float* coors_range_host,
                       int64_t num_points, int64_t num_features, int64_t max_points,
                       int64_t max_voxels, void* workspace, int64_t workspace_bytes,
                       void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  Grid3 g;
  DDF_CHECK_ARG(voxel_size_host && coors_range_host, "hard_voxelize: null voxel_size/range");
  int rc = make_grid(voxel_size_host, coors_range_host, &g);
  if (rc) return rc;
  DDF_CHECK_ARG(num_points >= 0 && num_features >= 3, "hard_voxelize: need (N, >=3) points");
  DDF_CHECK_ARG(max_points > 0 && max_voxels >= 0 && num_points < (1ll << 30),
                "hard_voxelize: bad max_points/max_voxels/num_points");
  DDF_CHECK_ARG(voxel_num != nullptr, "hard_voxelize: null voxel_num");
  if (num_points == 0 || max_voxels == 0) {
    DDF_CUDA(cudaMemsetAsync(voxel_num, 0, sizeof(int), stream));
    return DDF_OK;
  }
  DDF_CHECK_ARG(points && (voxels || mean) && coors && num_points_per_voxel, "hard_voxelize: null pointer");
  DDF_CHECK_ARG(voxels || (mean_features > 0 && mean_features <= num_features),
                "hard_voxelize_mean: need 0 < mean_features <= num_features");
  const int n = (int)num_points, F = (int)num_features, T = (int)max_points;
  VoxWs w = carve(workspace, n, T, (int)max_voxels);
  DDF_CHECK_ARG(workspace && (size_t)workspace_bytes >= w.bytes,
                "hard_voxelize: workspace too small (%lld < %lld)", (long long)workspace_bytes,
                (long long)w.bytes);
  const unsigned slots = table_slots(n);
  DDF_CUDA(cudaMemsetAsync(w.keys, 0xff, (size_t)slots * sizeof(int), stream));
  DDF_CUDA(cudaMemsetAsync(w.lists, 0x7f, (size_t)slots * T * sizeof(int), stream));
  const unsigned nb = (unsigned)ddf::cdiv(n, kThreads);
  DDF_LAUNCH(vox_insert_kernel, nb, kThreads, 0, stream, points, g, n, F, T, w.keys, slots - 1, w.lists,
                                                 w.slot);
  DDF_LAUNCH(vox_flag_first_kernel, nb, kThreads, 0, stream, w.slot, w.lists, T, n, w.is_first);
  rc = ddf::exclusive_scan_i32(w.is_first, w.rank, n, w.scan_ws, stream);
  if (rc) return rc;
  DDF_LAUNCH(vox_rank_kernel, nb, kThreads, 0, stream, w.is_first, w.rank, n, (int)max_voxels,
                                               w.first_point, w.cut, voxel_num);
  // upper bound on voxels = min(n, max_voxels); threads beyond the device-side voxel_num exit
  const long long vmax = n < max_voxels ? n : max_voxels;
  if (voxels) {
    const long long total = vmax * T * F;
    DDF_LAUNCH(vox_gather_kernel, (unsigned)ddf::cdiv(total, kThreads), kThreads, 0, stream,
        points, g, F, T, w.first_point, w.slot, w.keys, w.lists, w.cut, voxel_num, voxels, coors,
        num_points_per_voxel, total);
  } else {
    const long long total = vmax * mean_features;
    DDF_LAUNCH(vox_mean_kernel, (unsigned)ddf::cdiv(total, kThreads), kThreads, 0, stream,
        points, g, F, T, (int)mean_features, w.first_point, w.slot, w.keys, w.lists, w.cut, voxel_num, mean, coors,
        num_points_per_voxel, total);
  }
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}
}  // namespace

extern "C" int ddf_hard_voxelize(const float* points, float* voxels, int* coors,
                                 int* num_points_per_voxel, int* voxel_num,
                                 const float* voxel_size_host, const float* coors_range_host,
                                 int64_t num_points, int64_t num_features, int64_t max_points,
                                 int64_t max_voxels, void* workspace, int64_t workspace_bytes,
                                 void* stream_) {
  DDF_CHECK_ARG(num_points == 0 || max_voxels == 0 || voxels != nullptr, "hard_voxelize: null voxels");
  return hard_voxelize_impl(points, voxels, nullptr, 0, coors, num_points_per_voxel, voxel_num, voxel_size_host,
                            coors_range_host, num_points, num_features, max_points, max_voxels, workspace,
                            workspace_bytes, stream_);
}

// Hard voxelization fused with HardSimpleVFE: mean [max_voxels, mean_features] = mean over the (<= max_points)
// points of each voxel of the first mean_features point features; the padded voxel tensor is never written.
extern "C" int ddf_hard_voxelize_mean(const float* points, float* mean, int* coors, int* num_points_per_voxel,
                                      int* voxel_num, const float* voxel_size_host, const float* coors_range_host,
                                      int64_t num_points, int64_t num_features, int64_t mean_features,
                                      int64_t max_points, int64_t max_voxels, void* workspace,
                                      int64_t workspace_bytes, void* stream_) {
  DDF_CHECK_ARG(num_points == 0 || max_voxels == 0 || mean != nullptr, "hard_voxelize_mean: null mean");
  return hard_voxelize_impl(points, nullptr, mean, mean_features, coors, num_points_per_voxel, voxel_num,
                            voxel_size_host, coors_range_host, num_points, num_features, max_points, max_voxels,
                            workspace, workspace_bytes, stream_);
}
