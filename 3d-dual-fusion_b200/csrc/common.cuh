// Shared helpers for the ddf_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ddf_b200.h"

namespace ddf {

// thread-local last error string, surfaced through ddf_last_error()
void set_error(const char* fmt, ...);
const char* get_error();

constexpr int kNumSM = 148;  // B200

static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

// bookkeeping for bench.py's gpu_launches: every kernel launch of this library is counted
void note_launches(int n);

}  // namespace ddf

#define DDF_CHECK_ARG(cond, ...)   \
  do {                             \
    if (!(cond)) {                 \
      ddf::set_error(__VA_ARGS__); \
      return DDF_ERR_ARG;          \
    }                              \
  } while (0)

#define DDF_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      ddf::set_error("%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e__));   \
      return DDF_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

#define DDF_LAUNCH_CHECK() DDF_CUDA(cudaGetLastError())
// launch a kernel and count it:  DDF_LAUNCH(kernel<T>, grid, block, smem, stream, args...)
#define DDF_LAUNCH(kernel, grid, block, smem, stream, ...) \
  do {                                                     \
    kernel<<<grid, block, smem, stream>>>(__VA_ARGS__);    \
    ddf::note_launches(1);                                 \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per DEVICE for a kernel (the attribute belongs to
// the device's instance of the function; a process may drive several devices)
#define DDF_SET_SMEM_ONCE(kernel, bytes)                                                                 \
  do {                                                                                                   \
    static bool done__[64] = {};                                                                         \
    int dev__ = 0;                                                                                       \
    DDF_CUDA(cudaGetDevice(&dev__));                                                                     \
    if (dev__ < 0 || dev__ >= 64 || !done__[dev__]) {                                                    \
      DDF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      if (dev__ >= 0 && dev__ < 64) done__[dev__] = true;                                                \
    }                                                                                                    \
  } while (0)

// ---- small device helpers ------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

// vectorised fp32 reduction to global memory (sm_90+): one 16-byte red instead of four atomics
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b),
               "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
