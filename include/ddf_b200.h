/*
 * ddf_b200.h — C-ABI of the B200-native 3D-Dual-Fusion hot path (libddf_b200.so).
 *
 * Every entry point replaces one pybind11 / torch-extension function of the reference
 * (rasd3/3D-Dual-Fusion); the reference interface it stands in for is cited as file:line
 * relative to the reference root.  Conventions shared by all entry points:
 *
 *   - plain pointers + sizes only, no torch types; all data pointers are DEVICE pointers
 *     (sm_100a) unless the parameter name ends in _host;
 *   - the caller allocates every output (PyTorch's caching allocator on the Python side), the
 *     library never allocates device memory except where a `workspace` pointer is requested
 *     (size obtained from the matching *_workspace_bytes() call);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), no host synchronisation
 *     inside unless documented, re-entrant per stream;
 *   - return value: 0 = ok, non-zero = error code below; message via ddf_last_error()
 *     (thread-local).  The Python shims raise RuntimeError, mirroring the reference's
 *     AT_ASSERTM / TORCH_CHECK behaviour.
 */
#ifndef DDF_B200_H_
#define DDF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDF_OK 0
#define DDF_ERR_ARG 1
#define DDF_ERR_CUDA 2
#define DDF_ERR_UNSUPPORTED 3

/* ---- library ------------------------------------------------------------------------- */
const char* ddf_last_error(void);
/* ABI version, bumped whenever a signature changes. */
int ddf_abi_version(void);
/* Compiled SM architecture (100 for sm_100a). */
int ddf_compiled_arch(void);

/* ---- Multi-scale deformable attention (MSDA) ------------------------------------------
 * Replaces MultiScaleDeformableAttention.ms_deform_attn_forward / _backward
 *   reference: <proj>/models/model_utils/ops/src/ms_deform_attn.h:20-39, 41-62
 *              <proj>/models/model_utils/ops/src/cuda/ms_deform_attn_cuda.cu:20-80, 83-153
 *              kernels ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299 (fwd), 301-920 (bwd)
 *   Python caller: ops/functions/ms_deform_attn_func.py:21-38 (MSDeformAttnFunction)
 *
 * Layouts (all contiguous, as the reference asserts):
 *   value            [N, S, M, D]            (S = sum_l H_l*W_l)
 *   spatial_shapes   [L, 2]  int64 DEVICE    (H_l, W_l)
 *   level_start_idx  [L]     int64 DEVICE
 *   sampling_loc     [N, Lq, M, L, P, 2]     (x, y) normalised to [0,1] of the padded map
 *   attn_weight      [N, Lq, M, L, P]
 *   output           [N, Lq, M*D]            fully written (no pre-zeroing needed)
 *   grad_output      [N, Lq, M*D]
 *   grad_value       like value              zeroed inside, then accumulated with red.global
 *   grad_sampling_loc / grad_attn_weight     like sampling_loc / attn_weight, fully written
 * im2col_step is validated exactly like the reference (N % min(N, im2col_step) == 0) and is
 * otherwise unused: one launch covers the whole batch.
 * dtype: 0 = float32, 1 = float64 (the reference dispatches AT_DISPATCH_FLOATING_TYPES).
 */
int ddf_ms_deform_attn_forward(const void* value, const int64_t* spatial_shapes,
                               const int64_t* level_start_index, const void* sampling_loc,
                               const void* attn_weight, void* output, int64_t N, int64_t S,
                               int64_t M, int64_t D, int64_t L, int64_t Lq, int64_t P,
                               int64_t im2col_step, int dtype, void* stream);

int ddf_ms_deform_attn_backward(const void* value, const int64_t* spatial_shapes,
                                const int64_t* level_start_index, const void* sampling_loc,
                                const void* attn_weight, const void* grad_output,
                                void* grad_value, void* grad_sampling_loc,
                                void* grad_attn_weight, int64_t N, int64_t S, int64_t M,
                                int64_t D, int64_t L, int64_t Lq, int64_t P,
                                int64_t im2col_step, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DDF_B200_H_ */
