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

// programmatic dependent launch (DDF_LAUNCH_PDL); DDF_PDL=0 turns the launch attribute off (the kernels' griddepcontrol
// instructions are no-ops then)
int pdl_enabled();

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

// Same, with programmatic stream serialization: the grid may be scheduled while the previous kernel of the stream is
// still draining (its CTAs become resident as SMs free up and block in ddf::pdl_sync() until that kernel has completed
// and its writes are visible), which removes the ~2 us launch bubble between dependent kernels: 34.3 -> 33.9 ms per
// bench step for the conv / BatchNorm chain.  The extended launch costs the host ~2.5 us more than <<<>>>, so it is
// used where the GPU is the bottleneck (long back-to-back kernels) and NOT in the launch-bound parts of a step
// (voxelization, rule books, wrapper glue): with it on every launch the step was 1 ms slower, not faster.
// Every kernel of the library starts with ddf::pdl_sync() (tests/test_abi.py checks the SASS), so any launch site may be
// switched; DDF_PDL=0 turns the attribute off.
#define DDF_LAUNCH_PDL(kernel, grid_, block_, smem_, stream_arg_, ...)              \
  do {                                                                              \
    cudaLaunchConfig_t cfg__ = {};                                                  \
    cfg__.gridDim = dim3(grid_);                                                    \
    cfg__.blockDim = dim3(block_);                                                  \
    cfg__.dynamicSmemBytes = (size_t)(smem_);                                       \
    cfg__.stream = (stream_arg_);                                                   \
    cudaLaunchAttribute at__[1];                                                    \
    at__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                \
    at__[0].val.programmaticStreamSerializationAllowed = ddf::pdl_enabled();        \
    cfg__.attrs = at__;                                                             \
    cfg__.numAttrs = 1;                                                             \
    DDF_CUDA(cudaLaunchKernelEx(&cfg__, kernel, __VA_ARGS__));                      \
    ddf::note_launches(1);                                                          \
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
namespace ddf {
// First statement of EVERY kernel (so that any launch site may use DDF_LAUNCH_PDL): wait until the previous kernel of the stream has completed
// and flushed its writes, then let the NEXT kernel's CTAs be scheduled early in turn (they wait in their own pdl_sync).
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
}  // namespace ddf
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
