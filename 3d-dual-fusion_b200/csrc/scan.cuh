// Device-wide exclusive prefix sum over int32 (three short kernels, no host sync).
// Used to turn "is first point of its voxel" / "is new output cell" flags into dense ranks, which is
// what makes voxel order and rulebook order deterministic (first-seen / sorted) instead of
// atomic-arrival order as in the reference (voxelization_cuda.cu:150-180 does this scan on ONE
// thread; spconv's indice.cu.h:57,197 takes atomicAdd order).
#pragma once
#include "common.cuh"

namespace ddf {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;                              // per thread
constexpr int kScanTile = kScanThreads * kScanItems;       // 2048 per block

// bytes of scratch needed for n elements
static inline size_t scan_workspace_bytes(long long n) {
  long long nb = cdiv(n > 0 ? n : 1, kScanTile);
  return (size_t)(nb + 1) * sizeof(int);
}

// out[i] = sum_{j<i} in[j] for i in [0, n]; out has n+1 entries (out[n] = total). in may alias out
// only if in == out exactly is NOT supported (out is one longer) — use separate buffers.
int exclusive_scan_i32(const int* in, int* out, long long n, int* block_sums, cudaStream_t stream);

}  // namespace ddf
