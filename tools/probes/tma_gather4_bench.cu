// Issue-rate / bandwidth microbenchmark of TMA gather4 on sm_100a: W warps per CTA, L lanes per warp each
// issuing one gather4 (4 rows x 128 B) per iteration into a per-warp ring of 4 smem slots.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_gather4_bench tma_gather4_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>  // 0: L lanes issue in parallel (compiler waterfall); 1: lane 0 issues L gather4 back to back
__global__ void bench(const __grid_constant__ CUtensorMap tm, int n_rows, int L, int iters, unsigned* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[32 * 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = blockDim.x >> 5;
  if (lane == 0)
    for (int s = 0; s < 4; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[warp * 4 + s])));
  asm volatile("fence.mbarrier_init.release.cluster;");
  __syncthreads();
  uint8_t* my = smem + warp * (4 * L * 512);  // 4 slots x (L gather4 x 512 B)
  unsigned seed = (blockIdx.x * 977u + warp * 131u + lane * 7u) | 1u;
  for (int it = 0; it < iters; ++it) {
    const int s = it & 3;
    uint64_t* bar = &bars[warp * 4 + s];
    if (it >= 4) {  // slot reuse: wait for the load issued 4 iterations ago
      uint32_t done = 0;
      const uint32_t par = ((it >> 2) - 1) & 1;
      while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(s32(bar)), "r"(par));
    }
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(L * 512));
    __syncwarp();
    if (MODE == 0) {
      if (lane < L) {
        seed = seed * 1664525u + 1013904223u; const int r0 = (seed >> 8) % n_rows;
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                     ::"r"(s32(my + s * L * 512 + lane * 512)), "l"(&tm), "r"(s32(bar)), "r"(0), "r"(r0), "r"((r0 + 3) % n_rows), "r"((r0 + 11) % n_rows), "r"((r0 + 40) % n_rows) : "memory");
      }
    } else {
      if (lane == 0) {
#pragma unroll 4
        for (int j = 0; j < L; ++j) {
          seed = seed * 1664525u + 1013904223u; const int r0 = (seed >> 8) % n_rows;
          asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                       ::"r"(s32(my + s * L * 512 + j * 512)), "l"(&tm), "r"(s32(bar)), "r"(0), "r"(r0), "r"((r0 + 3) % n_rows), "r"((r0 + 11) % n_rows), "r"((r0 + 40) % n_rows) : "memory");
        }
      }
    }
  }
  // drain
  for (int it = (iters > 4 ? iters - 4 : 0); it < iters; ++it) {
    uint64_t* bar = &bars[warp * 4 + (it & 3)];
    uint32_t done = 0; const uint32_t par = (it >> 2) & 1;
    while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(s32(bar)), "r"(par));
  }
  if (threadIdx.x == 0) sink[blockIdx.x] = ((unsigned*)smem)[1] + W;
}

int main() {
  const int N = 131072, C = 128;  // 64 MiB, L2 resident
  float* d; unsigned* sink;
  cudaMalloc(&d, (size_t)N * C * 4); cudaMemset(d, 0, (size_t)N * C * 4); cudaMalloc(&sink, 4096);
  EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)N}, gstr[1] = {(cuuint64_t)C * 4};
  cuuint32_t box[2] = {32, 1}, estr[2] = {1, 1};
  enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int smem = 16 * 4 * 32 * 512 / 4 + 2048;  // sized per config below
  cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  (void)smem;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 400;
  int Ws[] = {1, 2, 4, 8, 16};
  int Ls[] = {1, 2, 4, 8, 32};
  for (int mode = 0; mode < 2; ++mode)
    for (int W : Ws)
      for (int L : Ls) {
        if ((size_t)W * 4 * L * 512 + 2048 > 200 * 1024) continue;
        size_t sm = (size_t)W * 4 * L * 512 + 2048;
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(a);
          if (mode == 0) bench<0><<<148, W * 32, sm>>>(tm, N, L, iters, sink); else bench<1><<<148, W * 32, sm>>>(tm, N, L, iters, sink);
          cudaEventRecord(b);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          float ms; cudaEventElapsedTime(&ms, a, b);
          if (rep == 1) {
            const double g4 = 148.0 * W * L * iters;
            printf("mode %d  warps %2d  lanes %2d : %.3f ms  %.1f clk/gather4/SM  %.2f TB/s\n", mode, W, L, ms,
                   ms * 1e-3 * 1.9e9 / (W * L * (double)iters), g4 * 512 / ms / 1e9);
          }
        }
      }
  return 0;
}
