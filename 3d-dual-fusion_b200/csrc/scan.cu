#include "scan.cuh"

namespace ddf {
namespace {

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total via smem
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
  __shared__ int wsum[kScanThreads / 32];
  __shared__ int tot;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int inc = warp_incl_scan(v, lane);
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = lane < kScanThreads / 32 ? wsum[lane] : 0;
    int si = warp_incl_scan(s, lane);
    if (lane < kScanThreads / 32) wsum[lane] = si - s;
    if (lane == kScanThreads / 32 - 1) tot = si;
  }
  __syncthreads();
  *total = tot;
  const int r = inc - v + wsum[w];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const int* __restrict__ in,
                                                                    long long n,
                                                                    int* __restrict__ block_sums) {
  ddf::pdl_sync();
  const long long base = (long long)blockIdx.x * kScanTile;
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const long long i = base + k * kScanThreads + threadIdx.x;
    if (i < n) s += in[i];
  }
  int tot;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

// single block: exclusive scan of the nb block sums in place, total -> block_sums[nb]
__global__ void __launch_bounds__(kScanThreads) scan_spine_kernel(int* block_sums, int nb) {
  ddf::pdl_sync();
  int carry = 0;
  for (int base = 0; base < nb; base += kScanThreads) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? block_sums[i] : 0;
    int tot;
    const int ex = block_excl_scan(v, &tot);
    if (i < nb) block_sums[i] = ex + carry;
    carry += tot;
  }
  if (threadIdx.x == 0) block_sums[nb] = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int* __restrict__ in,
                                                                   int* __restrict__ out,
                                                                   long long n,
                                                                   const int* __restrict__ block_sums,
                                                                   int nb) {
  ddf::pdl_sync();
  // thread owns kScanItems CONSECUTIVE elements so the in-thread order is the global order
  const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = base + k < n ? in[base + k] : 0;
    s += v[k];
  }
  int tot;
  int ex = block_excl_scan(s, &tot) + block_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = ex;
    ex += v[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = block_sums[nb];
}

}  // namespace

int exclusive_scan_i32(const int* in, int* out, long long n, int* block_sums, cudaStream_t stream) {
  const int nb = (int)cdiv(n > 0 ? n : 1, kScanTile);
  DDF_LAUNCH(scan_reduce_kernel, nb, kScanThreads, 0, stream, in, n, block_sums);
  DDF_LAUNCH(scan_spine_kernel, 1, kScanThreads, 0, stream, block_sums, nb);
  DDF_LAUNCH(scan_apply_kernel, nb, kScanThreads, 0, stream, in, out, n, block_sums, nb);
  DDF_LAUNCH_CHECK();
  return DDF_OK;
}

}  // namespace ddf
