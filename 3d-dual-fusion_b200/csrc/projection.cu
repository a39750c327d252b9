// LiDAR -> camera projection with camera assignment, and the stable per-(sample, camera) ranks of the zero-padded
// query layout — the int32 / geometry side of the TransFusion fusion wrapper as two kernels.
//
// Replaces get_2d_coor_multi + projection + the mask loops of split_param
//   TransFusion/mmdet3d/models/fusion_layers/point_fusion.py:509-549, 551-643, 342-382
// which run per sample and per camera on the HOST (NumPy quaternion chain through the nuScenes devkit, D2H copy at
// :585, boolean-mask loops with .sum() / .nonzero() syncs).  Here: one thread per voxel centre walks the cameras with
// the composed lidar2img matrices (explicit fp32 multiply-adds, never a tf32 GEMM: a pixel of error flips camera
// assignments and the // 4 feature pick), applies the reference's visibility rule and image transformation, and keeps
// the LAST camera that sees the voxel (unseen -> camera 0 at (0, 0)); a second kernel gives every query its stable rank
// inside its (sample, camera) group, which is the column of the padded layout (row order = input order, bit-identical
// to the reference's mask loops).
#include "common.cuh"

namespace {
constexpr int kThreads = 256;
constexpr int kMaxCam = 16;

struct ProjParams {
  float m[kMaxCam][12];       // rows 0..2 of lidar2img per camera
  float ori_w, ori_h;         // visibility test in ORIGINAL image pixels
  float sx, sy, cx, cy;       // scale, crop offset
  float flip_w;               // img_shape width when flipped, else < 0
  float pad_w, pad_h;         // padded input size (normaliser of grid)
  int n_cam;
};

__global__ void __launch_bounds__(kThreads)
project_assign_kernel(const float* __restrict__ pts, int stride, ProjParams P, int n, int group_base,
                      int* __restrict__ group, float* __restrict__ grid, float* __restrict__ grid_o) {
  ddf::pdl_sync();
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const float x = pts[(long long)i * stride], y = pts[(long long)i * stride + 1], z = pts[(long long)i * stride + 2];
  int cam = 0;
  float gx = 0.f, gy = 0.f;
  for (int c = 0; c < P.n_cam; ++c) {
    const float* m = P.m[c];
    // sum in the order of (l2i * homo).sum(-1): ((m0 x + m1 y) + m2 z) + m3, each product rounded (no FMA)
    const float u = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z)), m[3]);
    const float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[4], x), __fmul_rn(m[5], y)), __fmul_rn(m[6], z)), m[7]);
    const float d = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[8], x), __fmul_rn(m[9], y)), __fmul_rn(m[10], z)), m[11]);
    const float px = __fdiv_rn(u, d), py = __fdiv_rn(v, d);
    const bool seen = d > 1.0f && px > 1.f && px < P.ori_w - 1.f && py > 1.f && py < P.ori_h - 1.f;
    if (seen) {       // the last camera that sees the voxel wins
      cam = c;
      float qx = __fsub_rn(__fmul_rn(px, P.sx), P.cx);
      const float qy = __fsub_rn(__fmul_rn(py, P.sy), P.cy);
      if (P.flip_w >= 0.f) qx = __fsub_rn(P.flip_w, qx);
      gx = qx;
      gy = qy;
    }
  }
  group[i] = group_base + cam;
  grid_o[2 * i] = gx;
  grid_o[2 * i + 1] = gy;
  grid[2 * i] = __fdiv_rn(gx, P.pad_w);
  grid[2 * i + 1] = __fdiv_rn(gy, P.pad_h);
}

// one CTA per group: col[i] = number of j < i with group[j] == g (stable rank), counts[g] = group size
__global__ void __launch_bounds__(1024)
group_rank_kernel(const int* __restrict__ group, int n, int* __restrict__ col, int* __restrict__ counts) {
  ddf::pdl_sync();
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int start = 0; start < n; start += 1024) {
    const int i = start + threadIdx.x;
    const bool mine = i < n && group[i] == g;
    const unsigned b = __ballot_sync(0xffffffffu, mine);
    if (lane == 0) s_warp[warp] = __popc(b);
    __syncthreads();
    int before = 0;                                   // members in the warps before mine (32 adds; n / 1024 rounds)
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (mine) col[i] = s_base + before + __popc(b & ((1u << lane) - 1u));
    __syncthreads();
    if (threadIdx.x == 1023) s_base += before + __popc(b);
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[g] = s_base;
}
}  // namespace

// ---- CenterPoint / Det3D flavour: every camera that sees a voxel makes a query ---------------------------------
// Replaces Point2ImageProjection.transform_grid / forward of
//   CenterPoint/det3d/models/fusion/point_to_image_projection.py (+ voxel_with_point_projection.py:160-260): per
// camera, per-voxel GATHERED copies of the sample's 4x4 / 3x3 matrices, a chain of elementwise / reduction
// launches over (N, 4, 4) tensors and a rescale to feature-map pixels - 30 ms of a 105 ms CenterPoint step (torch
// profiler: 18 ms vectorized_gather + 11 ms index).  One thread per voxel walks the cameras; the arithmetic is the
// module's fp32 sequence (products rounded one by one, sums left to right, IEEE divisions, float -> int64 truncation).
namespace {
__global__ void __launch_bounds__(kThreads)
project_cameras_kernel(const int* __restrict__ indices, const float* __restrict__ pts, const float* __restrict__ l2c,
                       const float* __restrict__ intr, const float* __restrict__ shape,
                       const float* __restrict__ thres, int n_cam, int B, float image_scale, float Hf, float Wf, int n,
                       long long* __restrict__ grid, float* __restrict__ depth, bool* __restrict__ mask,
                       long long* __restrict__ fx, long long* __restrict__ fy) {
  ddf::pdl_sync();
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const int b = indices[4 * (long long)i];
  const float x = pts[3 * (long long)i], y = pts[3 * (long long)i + 1], z = pts[3 * (long long)i + 2];
  for (int c = 0; c < n_cam; ++c) {
    const float* m = l2c + ((long long)c * B + b) * 16;
    float cam[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      cam[j] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__ldg(m + 4 * j), x), __fmul_rn(__ldg(m + 4 * j + 1), y)),
                                   __fmul_rn(__ldg(m + 4 * j + 2), z)),
                         __ldg(m + 4 * j + 3));
    const float cx = __fdiv_rn(cam[0], cam[3]), cy = __fdiv_rn(cam[1], cam[3]), cz = __fdiv_rn(cam[2], cam[3]);
    const float* k = intr + ((long long)c * B + b) * 9;
    float im[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      im[j] = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(k + 3 * j), cx), __fmul_rn(__ldg(k + 3 * j + 1), cy)),
                        __fmul_rn(__ldg(k + 3 * j + 2), cz));
    long long gx = (long long)__fdiv_rn(im[0], im[2]), gy = (long long)__fdiv_rn(im[1], im[2]);   // .long()
    gx = (long long)__fmul_rn(image_scale, (float)gx);                                           // scaled-image pixels
    gy = (long long)__fmul_rn(image_scale, (float)gy);
    const float h = __ldg(shape + ((long long)c * B + b) * 2), w = __ldg(shape + ((long long)c * B + b) * 2 + 1);
    const bool ok = gx > 0 && (float)gx < w && gy > 0 && (float)gy < h && cz > __ldg(thres + c);
    const long long o = (long long)c * n + i;
    if (!ok) gx = gy = 0;
    grid[2 * o] = gx;
    grid[2 * o + 1] = gy;
    depth[o] = ok ? cz : 0.f;
    mask[o] = ok;
    // feature-map pixel: (grid.float() * (Wf / raw_w)).long()
    fx[o] = (long long)__fmul_rn((float)gx, __fdiv_rn(Wf, w));
    fy[o] = (long long)__fmul_rn((float)gy, __fdiv_rn(Hf, h));
  }
}
}  // namespace

// indices [n, 4] int32 (b, z, y, x); pts [n, 3] fp32; lidar2cam [n_cam, B, 4, 4], intrinsic [n_cam, B, 3, 3],
// image_shape [n_cam, B, 2] = (H, W) as fp32, depth_thres [n_cam]: device tensors.  Outputs, all [n_cam, n(, 2)]:
// grid int64 (x, y) in scaled-image pixels, depth, mask (bool), feat_x / feat_y int64 = feature-map pixel (Hf x Wf map).
extern "C" int ddf_project_cameras(const int* indices, const float* pts, const float* lidar2cam, const float* intrinsic,
                                   const float* image_shape, const float* depth_thres, int64_t n, int64_t n_cam,
                                   int64_t B, float image_scale, int64_t Hf, int64_t Wf, int64_t* grid, float* depth,
                                   void* mask, int64_t* feat_x, int64_t* feat_y, void* stream_) {
  DDF_CHECK_ARG(n >= 0 && n_cam > 0 && B > 0 && Hf > 0 && Wf > 0 && n < (1ll << 31), "project_cameras: bad sizes");
  if (n == 0) return DDF_OK;
  DDF_CHECK_ARG(indices && pts && lidar2cam && intrinsic && image_shape && depth_thres && grid && depth && mask &&
                    feat_x && feat_y,
                "project_cameras: null pointer");
  DDF_LAUNCH(project_cameras_kernel, (unsigned)ddf::cdiv(n, kThreads), kThreads, 0, (cudaStream_t)stream_, indices, pts,
             lidar2cam, intrinsic, image_shape, depth_thres, (int)n_cam, (int)B, image_scale, (float)Hf, (float)Wf, (int)n,
             reinterpret_cast<long long*>(grid), depth, reinterpret_cast<bool*>(mask),
             reinterpret_cast<long long*>(feat_x), reinterpret_cast<long long*>(feat_y));
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// points [n, stride >= 3] fp32 (xyz first); lidar2img_host [n_cam, 4, 4] HOST floats; group_base = sample * n_cam.
// Outputs: group [n] int32 = group_base + camera, grid [n, 2] = (x / pad_w, y / pad_h), grid_o [n, 2] padded-image pixels.
extern "C" int ddf_project_assign(const float* points, int64_t n, int64_t stride, const float* lidar2img_host,
                                  int64_t n_cam, float ori_h, float ori_w, float scale_x, float scale_y, float crop_x,
                                  float crop_y, float flip_w, float pad_h, float pad_w, int64_t group_base, int* group,
                                  float* grid, float* grid_o, void* stream_) {
  DDF_CHECK_ARG(n >= 0 && stride >= 3 && n_cam > 0 && n_cam <= kMaxCam, "project_assign: bad sizes (n_cam <= %d)", kMaxCam);
  if (n == 0) return DDF_OK;
  DDF_CHECK_ARG(points && lidar2img_host && group && grid && grid_o, "project_assign: null pointer");
  ProjParams P;
  for (int c = 0; c < n_cam; ++c)
    for (int k = 0; k < 12; ++k) P.m[c][k] = lidar2img_host[c * 16 + k];
  P.ori_w = ori_w; P.ori_h = ori_h; P.sx = scale_x; P.sy = scale_y; P.cx = crop_x; P.cy = crop_y;
  P.flip_w = flip_w; P.pad_w = pad_w; P.pad_h = pad_h; P.n_cam = (int)n_cam;
  DDF_LAUNCH(project_assign_kernel, (unsigned)ddf::cdiv(n, kThreads), kThreads, 0, (cudaStream_t)stream_, points,
             (int)stride, P, (int)n, (int)group_base, group, grid, grid_o);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

// group [n] int32 in [0, n_groups) -> col [n] (stable rank inside the group), counts [n_groups]
extern "C" int ddf_group_ranks(const int* group, int64_t n, int64_t n_groups, int* col, int* counts, void* stream_) {
  DDF_CHECK_ARG(n >= 0 && n_groups > 0 && n < (1ll << 31) && n_groups < 65536, "group_ranks: bad sizes");
  DDF_CHECK_ARG(counts != nullptr && (n == 0 || (group && col)), "group_ranks: null pointer");
  DDF_LAUNCH(group_rank_kernel, (unsigned)n_groups, 1024, 0, (cudaStream_t)stream_, group, (int)n, col, counts);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}
