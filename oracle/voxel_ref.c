/*
 * ORACLE (test infrastructure, NOT the product): plain-C restatement of the reference's hard /
 * dynamic voxelization, CPU path
 *   TransFusion/mmdet3d/ops/voxel/src/voxelization_cpu.cpp:8-41 (dynamic), :44-102 (hard)
 * including the `break` at the first point that would open voxel number max_voxels (:73).
 * The reference's dense coor_to_voxelidx[gz][gy][gx] lookup (:127-128, 340 MB for the nuScenes
 * grid) is replaced by an open-addressing map with the same key -> voxel index meaning.
 * Pinned by: the known-answer vector of TransFusion/tests/test_models/test_voxel_encoder/
 * test_voxel_generator.py:6-22, and the reference extension itself (oracle/_ref, when built).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int coord_of(const float* p, const float* vs, const float* rng, const int* grid, int* c /*z,y,x*/) {
  for (int j = 0; j < 3; ++j) {
    const int v = (int)floor((p[j] - rng[j]) / vs[j]); /* fp32 sub + div, as the reference */
    if (v < 0 || v >= grid[j]) return 0;
    c[2 - j] = v;
  }
  return 1;
}

static void grid_of(const float* vs, const float* rng, int* grid) {
  for (int i = 0; i < 3; ++i) grid[i] = (int)roundf((rng[3 + i] - rng[i]) / vs[i]);
}

void oracle_dynamic_voxelize(const float* points, int32_t* coors, const float* voxel_size,
                             const float* coors_range, int64_t n, int64_t F) {
  int grid[3];
  grid_of(voxel_size, coors_range, grid);
  for (int64_t i = 0; i < n; ++i) {
    int c[3];
    if (coord_of(points + i * F, voxel_size, coors_range, grid, c)) {
      coors[3 * i] = c[0]; coors[3 * i + 1] = c[1]; coors[3 * i + 2] = c[2];
    } else {
      coors[3 * i] = coors[3 * i + 1] = coors[3 * i + 2] = -1;
    }
  }
}

/* returns voxel_num; outputs must be zero-initialised by the caller like the reference's new_zeros */
int64_t oracle_hard_voxelize(const float* points, float* voxels, int32_t* coors, int32_t* num_points,
                             const float* voxel_size, const float* coors_range, int64_t n, int64_t F,
                             int64_t max_points, int64_t max_voxels) {
  int grid[3];
  grid_of(voxel_size, coors_range, grid);
  uint64_t slots = 1024;
  while (slots < (uint64_t)(2 * n + 2)) slots <<= 1;
  int64_t* mkey = (int64_t*)malloc(slots * sizeof(int64_t));
  int32_t* mval = (int32_t*)malloc(slots * sizeof(int32_t));
  for (uint64_t i = 0; i < slots; ++i) mkey[i] = -1;
  int64_t voxel_num = 0;
  for (int64_t i = 0; i < n; ++i) {
    int c[3];
    if (!coord_of(points + i * F, voxel_size, coors_range, grid, c)) continue;
    const int64_t key = ((int64_t)c[0] * grid[1] + c[1]) * grid[0] + c[2];
    uint64_t h = ((uint64_t)key * 0x9E3779B97F4A7C15ull) >> 20 & (slots - 1);
    while (mkey[h] != -1 && mkey[h] != key) h = (h + 1) & (slots - 1);
    int64_t voxelidx;
    if (mkey[h] == -1) {
      voxelidx = voxel_num;
      if (max_voxels != -1 && voxel_num >= max_voxels) break;
      voxel_num += 1;
      mkey[h] = key;
      mval[h] = (int32_t)voxelidx;
      for (int k = 0; k < 3; ++k) coors[3 * voxelidx + k] = c[k];
    } else {
      voxelidx = mval[h];
    }
    const int32_t num = num_points[voxelidx];
    if (max_points == -1 || num < max_points) {
      for (int64_t k = 0; k < F; ++k) voxels[(voxelidx * max_points + num) * F + k] = points[i * F + k];
      num_points[voxelidx] += 1;
    }
  }
  free(mkey);
  free(mval);
  return voxel_num;
}
