// Probe of cp.async.bulk.tensor.2d.tile::gather4 semantics on sm_100a (run on the GPU box):
//   * where the 4 gathered rows land in shared memory under SWIZZLE_128B (dst at +0 and +512 of a 1024-B atom)
//   * what a row index that is negative or >= the tensor height produces (expected: zero fill)
//   * whether the mbarrier transaction count is the full 4 x 128 B also for out-of-bounds rows
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_gather4_probe tma_gather4_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, int4 rows, int col, int dst_off, float* out, int* timed_out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar);
  const uint32_t sd = (uint32_t)__cvta_generic_to_shared(smem) + dst_off;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) ((float*)smem)[i] = -7.f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(512));
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(sd), "l"(&tm), "r"(sb), "r"(col), "r"(rows.x), "r"(rows.y), "r"(rows.z), "r"(rows.w) : "memory");
  }
  uint32_t done = 0;
  long long spins = 0;
  while (!done && spins < 20000000) {
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(sb));
    ++spins;
  }
  if (threadIdx.x == 0) *timed_out = done ? 0 : 1;
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = ((float*)smem)[i];
}

int main() {
  const int N = 64, C = 64;
  std::vector<float> h(N * C);
  for (int r = 0; r < N; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = r * 1000 + c;
  float *d, *out; int* to;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 512 * 4); cudaMalloc(&to, 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)N}, gstr[1] = {(cuuint64_t)C * 4};
  cuuint32_t box[2] = {32, 1}, estr[2] = {1, 1};
  CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)rc);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
  int4 cases[3] = {{5, 9, 2, 7}, {5, -1, 64, 7}, {-1, -1, 1000000, 63}};
  int offs[2] = {0, 512};
  std::vector<float> o(512);
  for (int ci = 0; ci < 3; ++ci)
    for (int oi = 0; oi < 2; ++oi) {
      probe<<<1, 128, 8192>>>(tm, cases[ci], 32, offs[oi], out, to);
      cudaError_t e = cudaDeviceSynchronize();
      int t; cudaMemcpy(&t, to, 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(o.data(), out, 2048, cudaMemcpyDeviceToHost);
      printf("case rows=(%d,%d,%d,%d) dst_off=%d err=%s timed_out=%d\n", cases[ci].x, cases[ci].y, cases[ci].z, cases[ci].w, offs[oi], cudaGetErrorString(e), t);
      int rows[4] = {cases[ci].x, cases[ci].y, cases[ci].z, cases[ci].w};
      int ok = 1;
      for (int r = 0; r < 4; ++r) {
        const int ar = offs[oi] / 128 + r;  // row inside the 1024-B atom
        printf("  row slot %d (src %d): ", r, rows[r]);
        for (int ch = 0; ch < 8; ++ch) {
          const int pos = (ar * 128 + ((ch ^ (ar & 7)) << 4)) / 4;
          const float got = o[pos];
          const bool inb = rows[r] >= 0 && rows[r] < N;
          const float want = inb ? rows[r] * 1000 + 32 + ch * 4 : 0.f;
          if (got != want) ok = 0;
          printf("%g ", got);
        }
        printf("\n");
      }
      printf("  swizzle+fill as expected: %s\n", ok ? "YES" : "NO");
    }
  return 0;
}
