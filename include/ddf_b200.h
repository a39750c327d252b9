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
/* Number of kernels this library has launched so far in this process (all threads); reset != 0
 * zeroes the counter after reading. Used by bench.py's gpu_launches. */
int64_t ddf_launch_count(int reset);

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

/* ---- tile-staged dual-query deformable attention (one level, 4 points, D in {8, 16}, fp32) --------------
 * Fuses what the reference module does around the op (ops/modules/ms_deform_attn.py:149-166: softmax over the
 * L*P logits, sampling_locations = reference_points + offsets / (W, H)) with the sampling kernel
 * (ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299) and its backward (:301-403).  Queries are binned by image
 * tile once per encoder forward (ddf_msda_plan; reference points are shared by all layers and by backward),
 * every CTA stages its tile + halo of `value` in shared memory with one 4-D TMA box load.
 *   value [N, H*W, M, D]; reference_points [NQ, 2]; offsets [NQ, M, 1, 4, 2] (raw Linear output, pixels);
 *   logits [NQ, M, 4]; out [NQ, M*D].  Backward returns grads wrt value, offsets and logits.
 * Queries: either the regular layout NQ = N * Lq (query i samples image i / Lq; query_batch = NULL), or a ragged list
 * of only the REAL queries of the zero-padded per-camera layout with query_batch [NQ] int32 = image of each query
 * (the row-wise work of the encoder then skips the padding).
 * ddf_msda_tile_supported: 1 when (M, D, L, P) is handled here (else use ddf_ms_deform_attn_*). */
int ddf_msda_tile_supported(int64_t M, int64_t D, int64_t L, int64_t P);
int64_t ddf_msda_plan_bytes(int64_t N, int64_t NQ, int64_t H, int64_t W);
int ddf_msda_plan(const float* reference_points, const int* query_batch, void* plan, int64_t N, int64_t NQ,
                  int64_t Lq, int64_t H, int64_t W, void* stream);
int ddf_msda_tile_forward(const float* value, const float* reference_points, const float* offsets,
                          const float* logits, const void* plan, float* out, int64_t N, int64_t H, int64_t W,
                          int64_t M, int64_t D, int64_t NQ, void* stream);
int ddf_msda_tile_backward(const float* value, const float* reference_points, const float* offsets,
                           const float* logits, const float* grad_out, const void* plan, float* grad_value,
                           float* grad_offsets, float* grad_logits, int64_t N, int64_t H, int64_t W, int64_t M,
                           int64_t D, int64_t NQ, void* stream);
int ddf_msda_tile_backward(const float* value, const float* reference_points, const float* offsets,
                           const float* logits, const float* grad_out, const void* plan, float* grad_value,
                           float* grad_offsets, float* grad_logits, int64_t N, int64_t H, int64_t W, int64_t M,
                           int64_t D, int64_t Lq, void* stream);

/* ---- Voxelization ---------------------------------------------------------------------
 * Replaces mmdet3d.ops.voxel.voxel_layer.hard_voxelize / dynamic_voxelize
 *   reference: TransFusion/mmdet3d/ops/voxel/src/voxelization.h:51-69 (hard), :71-86 (dynamic)
 *              GPU path voxelization_cuda.cu:184-326, CPU path voxelization_cpu.cpp:105-142
 *   Python caller: TransFusion/mmdet3d/ops/voxel/voxelize.py:41-58 (_Voxelization.forward)
 *
 *   points                [N, F] float32, F >= 3, xyz first
 *   voxels                [max_voxels, max_points, F] float32  rows [0, voxel_num) fully written
 *   coors                 [max_voxels, 3] int32 (z, y, x)      rows [0, voxel_num) written
 *   num_points_per_voxel  [max_voxels] int32                   rows [0, voxel_num) written
 *   voxel_num             [1] int32 DEVICE  (the reference returns it as a host int after a sync;
 *                         here the caller decides when to read it)
 *   voxel_size_host[3] (x,y,z), coors_range_host[6] (xyz min, xyz max): HOST float arrays
 *   workspace: device scratch of ddf_hard_voxelize_workspace_bytes(N, max_points, max_voxels) bytes
 * Voxel order = first-seen order of the points, in-voxel order = input order, processing stops at
 * the first point that would open voxel number max_voxels (bit-exact with the reference).
 * dynamic_voxelize writes (-1,-1,-1) for out-of-range points (CPU semantics, cpu.cpp:33-38).
 */
int64_t ddf_hard_voxelize_workspace_bytes(int64_t num_points, int64_t max_points,
                                          int64_t max_voxels);

int ddf_hard_voxelize(const float* points, float* voxels, int* coors, int* num_points_per_voxel,
                      int* voxel_num, const float* voxel_size_host, const float* coors_range_host,
                      int64_t num_points, int64_t num_features, int64_t max_points,
                      int64_t max_voxels, void* workspace, int64_t workspace_bytes, void* stream);

/* Hard voxelization fused with HardSimpleVFE (TransFusion/mmdet3d/models/voxel_encoders/voxel_encoder.py:27-44):
 * mean [max_voxels, mean_features] float32 = sum of the first mean_features features over the voxel's points (list
 * order) / num_points; same coors / num_points / voxel_num / order contract as ddf_hard_voxelize, the padded
 * [max_voxels, max_points, F] tensor is never materialised. */
int ddf_hard_voxelize_mean(const float* points, float* mean, int* coors, int* num_points_per_voxel, int* voxel_num,
                           const float* voxel_size_host, const float* coors_range_host, int64_t num_points,
                           int64_t num_features, int64_t mean_features, int64_t max_points, int64_t max_voxels,
                           void* workspace, int64_t workspace_bytes, void* stream);

int ddf_dynamic_voxelize(const float* points, int* coors, const float* voxel_size_host,
                         const float* coors_range_host, int64_t num_points, int64_t num_features,
                         void* stream);

/* ---- Sparse-convolution rulebook -------------------------------------------------------
 * Replaces sparse_conv_ext.get_indice_pairs_3d
 *   reference: TransFusion/mmdet3d/ops/spconv/include/spconv/spconv_ops.h:27-141 (getIndicePair<3>),
 *              kernels include/spconv/indice.cu.h:24-234, CPU semantics include/spconv/geometry.h:24-297
 *   Python caller: TransFusion/mmdet3d/ops/spconv/ops.py:46-105 (get_indice_pairs)
 *
 *   indices        [N, 4] int32 (b, z, y, x), 16-byte aligned
 *   geometry       HOST int64[3] arrays (z, y, x order) exactly as the reference passes them; for
 *                  SubM stride is forced to 1 and padding to ksize/2 (spconv_ops.h:76-79)
 *   indice_pairs   [K, 2, N] int32: row 0 = input rows (ASCENDING inside an offset), row 1 = output
 *                  rows, tail filled with -1;  indice_num [K] int32
 *   out_indices    [num_act_out, 4] int32, sorted by flat (b,z,y,x) (reference GPU order)
 *   gather_table   optional [N_out, K] int32: input row feeding output o through offset k, or -1
 *   scatter_table  optional [N_in,  K] int32: output row fed by input j through offset k, or -1
 * Regular convs are two calls on the same stream + workspace: ddf_conv_count_outputs (candidate
 * cells -> bitmap -> count, written to a DEVICE int), then — after the caller sized its outputs —
 * ddf_conv_indice_pairs.  Transposed (deconv) rulebooks are not generated (no 3D-DF backbone uses
 * them); SparseInverseConv reuses a saved rulebook with inverse=1 below.
 * ddf_subm_indice_pairs: indice_pairs (and indice_num) may be NULL to build the tables only.
 */
int64_t ddf_indice_pairs_workspace_bytes(int64_t num_in, int64_t batch_size,
                                         const int64_t* out_spatial_shape,
                                         const int64_t* spatial_shape, const int64_t* ksize,
                                         const int64_t* stride, const int64_t* padding,
                                         const int64_t* dilation, int subm);

int ddf_subm_indice_pairs(const int* indices, int64_t num_in, int64_t batch_size,
                          const int64_t* spatial_shape, const int64_t* ksize,
                          const int64_t* dilation, int* indice_pairs, int* indice_num,
                          int* gather_table, int* scatter_table, void* workspace,
                          int64_t workspace_bytes, void* stream);

int ddf_conv_count_outputs(const int* indices, int64_t num_in, int64_t batch_size,
                           const int64_t* out_spatial_shape, const int64_t* spatial_shape,
                           const int64_t* ksize, const int64_t* stride, const int64_t* padding,
                           const int64_t* dilation, int* num_act_out, void* workspace,
                           int64_t workspace_bytes, void* stream);

int ddf_conv_indice_pairs(const int* indices, int64_t num_in, int64_t batch_size,
                          const int64_t* out_spatial_shape, const int64_t* spatial_shape,
                          const int64_t* ksize, const int64_t* stride, const int64_t* padding,
                          const int64_t* dilation, int64_t num_act_out, int* out_indices,
                          int* indice_pairs, int* indice_num, int* gather_table,
                          int* scatter_table, void* workspace, int64_t workspace_bytes,
                          void* stream);

/* ---- Sparse convolution ----------------------------------------------------------------
 * Replaces sparse_conv_ext.indice_conv_fp32 / indice_conv_backward_fp32
 *   reference: spconv_ops.h:260-361 (indiceConv), :363-456 (indiceConvBackward)
 *   Python callers: TransFusion/mmdet3d/ops/spconv/functional.py:20-98
 * features [N_in, Cin], filters [K, Cin, Cout] (= the reference's [kd,kh,kw,Cin,Cout] viewed flat),
 * out [N_out, Cout]; all float32, fully overwritten.
 * Drop-in entry points (ddf_indice_conv*) take the reference-format rulebook and the same
 * inverse / subm flags; scratch: table_ws = int32 [rows, K], filters_t_ws = float [K*Cin*Cout].
 * Fast entry points (ddf_sparse_conv_*) take the row-major tables of the rulebook build directly.
 */
int ddf_indice_conv(const float* features, const float* filters, const int* indice_pairs,
                    const int* indice_num, int64_t pair_stride, float* out, int64_t n_out,
                    int64_t kvol, int64_t cin, int64_t cout, int inverse, int subm, int* table_ws,
                    void* stream);

int ddf_indice_conv_backward(const float* features, const float* filters, const float* grad_out,
                             const int* indice_pairs, const int* indice_num, int64_t pair_stride,
                             float* grad_in, float* grad_filters, int64_t n_in, int64_t kvol,
                             int64_t cin, int64_t cout, int inverse, int subm, int* table_ws,
                             float* filters_t_ws, void* stream);

/* filters_t_ws: optional float [K*Cin*Cout] scratch; when given (and Cin % 32 == 0, Cout <= 128,
 * K <= 27) a tcgen05 implicit-GEMM kernel runs, otherwise the fp32 SIMT kernel.
 * operand_format 0: features are fp32 rows (tcgen05 kind::tf32). 1: features are in the bf16 hi/lo block
 * layout written by ddf_split_bf16x3 ("bf16x3": three bf16 MMAs per product pair, 16-bit significand,
 * fp32 accumulation); only where ddf_sparse_conv_tc_mode reports bit 4 (forward) / bit 5 (dgrad). */
int ddf_sparse_conv_forward(const float* features, const float* filters, const int* gather_table,
                            const float* bias, float* out, float* filters_t_ws, int64_t n_out,
                            int64_t n_in, int64_t kvol, int64_t cin, int64_t cout, int operand_format,
                            void* stream);

/* Tensor-core bookkeeping: ddf_sparse_conv_tc_mode returns a bit mask of the kernels of a layer
 * that run as tcgen05 implicit GEMMs (1 forward, 2 dgrad, 4 wgrad, 8 table-driven wgrad
 * available, 16 / 32 forward / dgrad expect operand_format 1; 0 with DDF_DISABLE_TC=1).
 * tf32 keeps 10 mantissa bits and the hardware TRUNCATES fp32 operands; ddf_round_tf32 rounds a
 * tensor to the nearest tf32 first (dst may alias src) so the error is unbiased. Filters are
 * rounded inside the conv calls. */
int ddf_sparse_conv_tc_mode(int64_t kvol, int64_t cin, int64_t cout);
/* Conv kernel selection: 4 (default; env DDF_TC_MODE overrides) as 1, with forward and dgrad of the layers
 * the multi-tile kernel takes in bf16x3 (wgrad stays tf32); 1 tcgen05 tf32, multi-tile kernel (filter slices by tiled TMA,
 * shared by up to 4 row tiles; rows gathered by cp.async) where the layer shape allows, else the
 * single-tile cp.async kernel; 2 single-tile cp.async kernel only; 3 as 1 with rows gathered by TMA
 * gather4; 0 fp32 SIMT kernels (full fp32 products). Returns the previous setting. Not thread-safe
 * against concurrent conv calls. */
int ddf_set_tensor_cores(int on);
int ddf_round_tf32(const float* src, float* dst, int64_t n, void* stream);
/* src fp32 [rows, cols], cols % 32 == 0 -> split [rows*cols*4 bytes]: per row, per 32 channels, 128 bytes
 * [32 x bf16 hi | 32 x bf16 lo], hi = bf16(x), lo = bf16(x - hi); rounded (may be NULL): tf32-rounded copy. */
int ddf_split_bf16x3(const float* src, void* split, float* rounded, int64_t rows, int64_t cols, void* stream);

/* wgrad through the forward gather table [n_out, K] instead of the pair lists (tcgen05; Cin, Cout in
 * {32, 64, 128}, K <= 27): output rows walked once, grad_out read densely. grad_filters is zeroed
 * inside. ddf_sparse_conv_tc_mode bit 3 (8) says whether a layer shape is supported. */
int ddf_sparse_conv_wgrad_table(const float* features, const float* grad_out, const int* gather_table,
                                float* grad_filters, int64_t n_out, int64_t n_in, int64_t kvol,
                                int64_t cin, int64_t cout, void* stream);

/* n_out = rows of grad_out (-1 when unknown: the TMA-staged kernel, whose tensor map needs the
 * height of the gathered tensor, is then not used). operand_format as in ddf_sparse_conv_forward
 * (1: grad_out is in the bf16 hi/lo block layout). */
int ddf_sparse_conv_dgrad(const float* grad_out, const float* filters, const int* scatter_table,
                          float* grad_in, float* filters_t_ws, int64_t n_in, int64_t n_out, int64_t kvol,
                          int64_t cin, int64_t cout, int operand_format, void* stream);

int ddf_sparse_conv_wgrad(const float* features, const float* grad_out, const int* indice_pairs,
                          const int* indice_num, int64_t pair_stride, float* grad_filters,
                          int64_t kvol, int64_t cin, int64_t cout, int inverse, void* stream);

/* ---- dense(): sparse -> dense NCDHW ----------------------------------------------------
 * Replaces SparseConvTensor.dense() / scatter_nd (TransFusion/mmdet3d/ops/spconv/structure.py:5-18,
 * 55-64): out [B, C, D, H, W] zeroed inside then written directly in NCDHW (no permute copy);
 * ddf_dense_to_sparse is its backward (gather of grad_dense at the active cells).
 */
int ddf_sparse_to_dense(const float* features, const int* indices, float* out, int64_t n,
                        int64_t C, int64_t B, int64_t D, int64_t H, int64_t W, void* stream);

int ddf_dense_to_sparse(const float* grad_dense, const int* indices, float* grad_features,
                        int64_t n, int64_t C, int64_t B, int64_t D, int64_t H, int64_t W,
                        void* stream);

/* BEV hand-off to a bf16 channels-last 2-D backbone (the step after the path: sparse_encoder.py:366-367 ->
 * backbones/second.py): out = bf16 [B, H, W, C*D], the channels-last storage of the logical [B, C*D, H, W] map
 * (channel c*D + z), zeroed inside; the backward gathers a bf16 gradient of the same layout into fp32 rows. */
int ddf_sparse_to_bev_nhwc_bf16(const float* features, const int* indices, void* out, int64_t n, int64_t C,
                                int64_t B, int64_t D, int64_t H, int64_t W, void* stream);
int ddf_bev_nhwc_bf16_to_sparse(const void* grad_dense, const int* indices, float* grad_features, int64_t n,
                                int64_t C, int64_t B, int64_t D, int64_t H, int64_t W, void* stream);

/* ---- Point-set ops of the 3D local self-attention (LocalTransformer) ---------------------
 * Replace furthest_point_sample_ext / ball_query_ext / group_points_ext / gather_points_ext
 *   reference: <proj>/ops/furthest_point_sample/src/furthest_point_sample.cpp (wrapper),
 *              furthest_point_sample_cuda.cu:25-141; <proj>/ops/ball_query/src/ball_query.cpp:30-43,
 *              ball_query_cuda.cu:11-54; <proj>/ops/group_points/src/group_points_cuda.cu:10-79;
 *              <proj>/ops/gather_points/src/gather_points_cuda.cu:8-70
 *   Python callers: furthest_point_sample.py:7-40, ball_query.py:7-47, group_points.py:153-208,
 *              gather_points.py:7-52 (<proj> = TransFusion/mmdet3d, CenterPoint/det3d, ...)
 * All tensors float32 / int32, contiguous.
 *   furthest_point_sampling: xyz [B,N,3], temp [B,N] running min distance (in/out; NULL = 1e10
 *       start, not written), idx [B,m] out. Picks are bit-identical to the reference kernel incl.
 *       its tie-break (see pointops.cu). N <= 65536.
 *   ball_query: new_xyz [B,m,3], xyz [B,N,3], idx [B,m,nsample] (ZERO-initialised by the caller as
 *       in ball_query.py:36): first nsample points in index order with d2 == 0 or
 *       min_r^2 <= d2 < max_r^2; unfilled slots repeat the first hit.
 *   group_points: out[b,c,p,s] = features[b,c,idx[b,p,s]]; *_grad scatter-adds (grad zeroed inside).
 *   gather_points: out[b,c,p] = points[b,c,idx[b,p]]; *_grad likewise.
 */
int ddf_furthest_point_sampling(const float* xyz, float* temp, int* idx, int64_t B, int64_t N,
                                int64_t m, void* stream);
int ddf_ball_query(const float* new_xyz, const float* xyz, int* idx, int64_t B, int64_t N, int64_t m,
                   float min_radius, float max_radius, int64_t nsample, void* stream);
int ddf_group_points(const float* features, const int* idx, float* out, int64_t B, int64_t C,
                     int64_t N, int64_t npoints, int64_t nsample, void* stream);
int ddf_group_points_grad(const float* grad_out, const int* idx, float* grad_features, int64_t B,
                          int64_t C, int64_t N, int64_t npoints, int64_t nsample, void* stream);
int ddf_gather_points(const float* points, const int* idx, float* out, int64_t B, int64_t C,
                      int64_t N, int64_t npoints, void* stream);
int ddf_gather_points_grad(const float* grad_out, const int* idx, float* grad_points, int64_t B,
                           int64_t C, int64_t N, int64_t npoints, void* stream);

/* LocalTransformer.scatter, 'unique' aggregation (<proj>/models/model_utils/pointformer.py:319-347): every voxel
 * takes the grouped feature of its FIRST occurrence in flattened (group, slot) order.
 *   ddf_first_occurrence: idx [B, E] (E = npoint*nsample, values in [0, N)) -> first [B, N] (E where never hit)
 *   ddf_scatter_first:    inout [B, C, N]: hit columns replaced by feats [B, C, E][:, first]
 *   ddf_scatter_first_grad: grad_feats [B, C, E] (zeroed inside), grad_features [B, C, N] (0 at hit columns) */
int ddf_first_occurrence(const int* idx, int* first, int64_t B, int64_t N, int64_t E, void* stream);
int ddf_scatter_first(const float* feats, const int* first, float* inout, int64_t B, int64_t C, int64_t N,
                      int64_t E, void* stream);
int ddf_scatter_first_grad(const float* grad_out, const int* first, float* grad_feats, float* grad_features,
                           int64_t B, int64_t C, int64_t N, int64_t E, void* stream);

/* 3D local self-attention core (LocalTransformer; <proj>/models/model_utils/pointformer.py:10-44, 349-380): inside
 * every ball-query group of group_size = 32 tokens, per head: out = softmax(q k^T / sqrt(head_dim)) v.
 *   qkv [groups*32, 3*heads*head_dim] float32 token-major (q | k | v as nn.MultiheadAttention's in_proj emits them)
 *   out [groups*32, heads*head_dim];  backward recomputes the probabilities: grad_qkv from grad_out.
 * head_dim in {16, 32}; ddf_local_attn_supported says whether a shape is taken. */
int ddf_local_attn_supported(int64_t heads, int64_t head_dim, int64_t group_size);
int ddf_local_attn_forward(const float* qkv, float* out, int64_t groups, int64_t heads, int64_t head_dim,
                           int64_t group_size, void* stream);
int ddf_local_attn_backward(const float* qkv, const float* grad_out, float* grad_qkv, int64_t groups, int64_t heads,
                            int64_t head_dim, int64_t group_size, void* stream);

/* ---- BatchNorm1d over sparse-voxel features [N, C], fused with the residual add and ReLU -----------
 * Replaces the elementwise chain conv -> BN1d -> [+ identity] -> ReLU of the reference's sparse
 * blocks (TransFusion/mmdet3d/ops/sparse_block.py:102-120,153-185; torch.nn.BatchNorm1d semantics:
 * batch statistics with biased variance in training, running statistics updated with the unbiased
 * variance and `momentum`; running statistics in eval).  C: power of two in [4, 1024]; fp32; all
 * pointers 16-byte aligned; residual / weight / bias / running_* may be NULL.
 * workspace: ddf_sparse_bn_workspace_bytes(C) bytes, ZERO-INITIALISED ONCE by the caller and then
 * reusable by successive calls on the same stream (the kernels leave it ready for the next call). */
int64_t ddf_sparse_bn_workspace_bytes(int64_t C);
int ddf_sparse_bn_forward(const float* x, const float* residual, const float* weight, const float* bias,
                          float* running_mean, float* running_var, float* y, float* save_mean,
                          float* save_invstd, int64_t n, int64_t C, int training, float momentum,
                          float eps, int relu, void* workspace, void* stream);
/* As ddf_sparse_bn_forward; the same pass also writes y in the bf16x3 operand layout of the tcgen05 convs (split: same
 * bytes as y, [32 x bf16 hi | 32 x bf16 lo] per 32 channels as ddf_split_bf16x3; C % 32 == 0) and, when rounded != NULL,
 * the tf32-rounded copy of y the wgrad kernels read. */
int ddf_sparse_bn_forward_split(const float* x, const float* residual, const float* weight, const float* bias,
                                float* running_mean, float* running_var, float* y, float* save_mean,
                                float* save_invstd, void* split, float* rounded, int64_t n, int64_t C, int training,
                                float momentum, float eps, int relu, void* workspace, void* stream);
/* mean / invstd: what forward normalised with (saved batch statistics in training; running_mean and
 * 1/sqrt(running_var + eps) in eval).  y is read only for the ReLU mask.  grad_x / grad_residual /
 * grad_weight / grad_bias may be NULL. */
int ddf_sparse_bn_backward(const float* grad_y, const float* y, const float* x, const float* weight,
                           const float* mean, const float* invstd, float* grad_x, float* grad_residual,
                           float* grad_weight, float* grad_bias, int64_t n, int64_t C, int training,
                           int relu, void* workspace, void* stream);

/* ---- Fused elementwise kernels of the 3D-DF encoder layers ---------------------------------------
 * Replace chains of PyTorch kernels in <proj>/models/model_utils/actr_transformer.py:383-397 (FFN:
 * linear2(dropout(relu(linear1(x))))) and :385,391,406 (norm(src + dropout(src2))).  fp32, 16-byte
 * aligned pointers.  Dropout is a counter-based hash of (seed, element index): p = drop probability
 * (0 in eval mode), kept values are scaled by 1/(1-p), nothing is stored for backward.
 *   ddf_bias_relu_dropout_forward : out[n, C] = dropout(relu(h + bias)); out may alias h; C % 4 == 0
 *   ddf_bias_relu_dropout_backward: grad_h = grad_out * (out != 0) / (1 - p); grad_bias [C] (optional, C/4
 *       must divide 256) is ACCUMULATED into: column sums of grad_h in the same pass
 *   ddf_add_dropout_layer_norm_forward : s = a + dropout(b) (b may be NULL), y = LayerNorm(s) * gamma
 *       + beta over the last dim C in {128, 256, 512}; s_out (optional), mean / rstd [rows] saved
 *   ddf_add_dropout_layer_norm_backward: grad_a, grad_b (either may be NULL); grad_gamma / grad_beta [C]
 *       are ACCUMULATED into (the caller zeroes them) */
int ddf_bias_relu_dropout_forward(const float* h, const float* bias, float* out, int64_t n, int64_t C,
                                  float p, uint64_t seed, void* stream);
int ddf_bias_relu_dropout_backward(const float* grad_out, const float* out, float* grad_h, float* grad_bias,
                                   int64_t n, int64_t C, float p, void* stream);
int ddf_add_dropout_layer_norm_forward(const float* a, const float* b, const float* gamma, const float* beta,
                                       float* s_out, float* y, float* mean, float* rstd, int64_t rows,
                                       int64_t C, float p, uint64_t seed, float eps, void* stream);
int ddf_add_dropout_layer_norm_backward(const float* grad_y, const float* s, const float* gamma,
                                        const float* mean, const float* rstd, float* grad_a, float* grad_b,
                                        float* grad_gamma, float* grad_beta, int64_t rows, int64_t C,
                                        float p, uint64_t seed, void* stream);

/* out [C] = column sums of x [rows, C] (the bias gradient of a Linear over all tokens; zeroed inside).
 * C % 4 == 0, C <= 1024. */
int ddf_col_sum(const float* x, float* out, int64_t rows, int64_t C, void* stream);

/* Bi-directional gated fusion of the LiDAR / image query streams (<proj>/models/model_utils/attentions.py:89-117,
 * BiGateSum1D and BiGateSum1D_2): o1 = f1 + f2 * s1, o2 = f2 + f1 * s2 with s1 = sigmoid(wb . u1 + bb),
 * s2 = sigmoid(wa . u2 + ba); fuse_in: u1 = u2 = f1 + f2 (BiGateSum1D_2), else u1 = f1, u2 = f2.  f*, o* [rows, C]
 * fp32 (C = 128 or 256), wb / wa [C] = the Conv1d(C, 1, 1) weights, bb / ba [1] (may be NULL), gates [rows, 2] is
 * kept for backward.  Backward: grad_o1 / grad_o2 / grad_f1 / grad_f2 / the parameter gradients may be NULL. */
int ddf_bigate_sum_forward(const float* f1, const float* f2, const float* wb, const float* bb, const float* wa,
                           const float* ba, float* o1, float* o2, float* gates, int64_t rows, int64_t C,
                           int fuse_in, void* stream);
int ddf_bigate_sum_backward(const float* grad_o1, const float* grad_o2, const float* f1, const float* f2,
                            const float* gates, const float* wb, const float* wa, float* grad_f1, float* grad_f2,
                            float* grad_wb, float* grad_bb, float* grad_wa, float* grad_ba, int64_t rows, int64_t C,
                            int fuse_in, void* stream);

/* FFN forward of the encoder layers as one kernel (<proj>/models/model_utils/actr_transformer.py:383-397):
 * h [T, F] = dropout(relu(x [T, D] . w1 [F, D]^T + b1)) (kept for backward), y [T, D] = h . w2 [D, F]^T + b2.
 * D = 128, F % 64 == 0 (ddf_ffn_supported); fp32 row-major, tf32 products, fp32 accumulation; dropout by a counter
 * hash of (seed, element index) - nothing is stored, ddf_bias_relu_dropout_backward applies to h unchanged (it reads
 * the pattern off h != 0). */
int ddf_ffn_supported(int64_t T, int64_t D, int64_t F);
int64_t ddf_ffn_workspace_bytes(int64_t D, int64_t F);   /* scratch for the re-laid weights, any content */
int ddf_ffn_forward(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, float* h,
                    float* y, void* workspace, int64_t T, int64_t D, int64_t F, float p, uint64_t seed, void* stream);
/* the drop probability ddf_ffn_forward really applies for a requested p (its threshold is quantised to 1 / 256):
 * pass THIS to ddf_bias_relu_dropout_backward, which scales by 1 / (1 - p) */
float ddf_ffn_dropout_p(float p);

/* ---- camera-side input projection in the token-major ("rows") layout -------------------------------------------
 * The reference runs Conv2d(k=1) + GroupNorm(32, d_model) on NCHW camera maps and flattens / transposes the result to
 * [B', H*W, d_model] (<proj>/models/model_utils/actr.py:131-187, actr_transformer.py:255-264); the per-query camera
 * feature takes Conv1d(k=1) + GroupNorm between two transposes (actr.py:150-158).  Here the map becomes rows once, the
 * 1x1 convolution is a row-major GEMM and GroupNorm runs on rows (same arithmetic as torch.nn.GroupNorm: biased
 * variance over the (L, C / G) elements of a (sample, group)).
 * ddf_nchw_to_rows: src [N, C, HW] (dtype 0 = fp32, 1 = bf16) -> dst [N, HW, C] fp32.
 * ddf_group_norm_rows_*: x / y [N, L, C], C == 4 * G (ddf_group_norm_rows_supported), w / b [C] may be NULL, mean / rstd
 * [N, G] saved by forward for backward, ws = N * G * 2 doubles of scratch (any content); backward overwrites grad_w /
 * grad_b [C] (may be NULL). */
int ddf_group_norm_rows_supported(int64_t C, int64_t G);
int ddf_nchw_to_rows(const void* src, int dtype, float* dst, int64_t N, int64_t C, int64_t HW, void* stream);
int ddf_group_norm_rows_forward(const float* x, const float* w, const float* b, float* y, float* mean, float* rstd,
                                void* ws, int64_t N, int64_t L, int64_t C, int64_t G, float eps, void* stream);
int ddf_group_norm_rows_backward(const float* grad_y, const float* x, const float* w, const float* mean,
                                 const float* rstd, float* grad_x, float* grad_w, float* grad_b, void* ws, int64_t N,
                                 int64_t L, int64_t C, int64_t G, void* stream);

/* c [M, N] = a [K, M]^T . b [K, N]: the weight gradient of an nn.Linear over K tokens, W.grad [out, in] =
 * grad_out [K, out]^T . x [K, in] (autograd of F.linear in <proj>/models/model_utils/actr_transformer.py:383-397,
 * ops/modules/ms_deform_attn.py:124-147).  fp32 row-major operands read as tf32 (top 19 bits) by tcgen05, fp32
 * accumulation, split-K partial sums reduced in c (overwritten).  M, N multiples of 32 (ddf_xty_supported). */
int ddf_xty_supported(int64_t K, int64_t M, int64_t N);
int ddf_xty_tf32(const float* a, const float* b, float* c, int64_t K, int64_t M, int64_t N, void* stream);

/* ---- fusion wrapper geometry (TransFusion/mmdet3d/models/fusion_layers/point_fusion.py:342-382, 509-643) ----------
 * ddf_project_assign: voxel centres -> camera assignment ("last camera that sees the voxel", unseen -> camera 0 at
 *   (0, 0)) + reference points; visibility in ORIGINAL image pixels (depth > 1, 1 < u < ori_w - 1, 1 < v < ori_h - 1),
 *   then scale -> crop -> flip (flip_w = un-padded image width, < 0 when not flipped) -> / padded size.
 *   points [n, stride] device fp32; lidar2img_host [n_cam, 4, 4] HOST floats (composed lidar -> pixel matrices);
 *   group [n] int32 = group_base + camera; grid [n, 2] normalised; grid_o [n, 2] padded-image pixels.
 * ddf_group_ranks: col [n] = stable rank of every element inside its group (= its column in the zero-padded
 *   (n_groups, max_count) layout, input order preserved), counts [n_groups]. */
int ddf_project_assign(const float* points, int64_t n, int64_t stride, const float* lidar2img_host, int64_t n_cam,
                       float ori_h, float ori_w, float scale_x, float scale_y, float crop_x, float crop_y, float flip_w,
                       float pad_h, float pad_w, int64_t group_base, int* group, float* grid, float* grid_o,
                       void* stream);
int ddf_group_ranks(const int* group, int64_t n, int64_t n_groups, int* col, int* counts, void* stream);

/* CenterPoint / Det3D projection (CenterPoint/det3d/models/fusion/point_to_image_projection.py transform_grid +
 * forward, voxel_with_point_projection.py:160-260): every camera that sees a voxel makes a query.  indices [n, 4]
 * int32 (b, z, y, x), pts [n, 3] reverse-augmented voxel corners, lidar2cam [n_cam, B, 4, 4], intrinsic
 * [n_cam, B, 3, 3], image_shape [n_cam, B, 2] (H, W) fp32, depth_thres [n_cam] - device tensors.  Outputs per
 * (camera, voxel): grid int64 [n_cam, n, 2] (x, y) in scaled-image pixels (0 where masked), depth, mask (bool bytes),
 * feat_x / feat_y int64 = pixel of the Hf x Wf feature map. */
int ddf_project_cameras(const int* indices, const float* pts, const float* lidar2cam, const float* intrinsic,
                        const float* image_shape, const float* depth_thres, int64_t n, int64_t n_cam, int64_t B,
                        float image_scale, int64_t Hf, int64_t Wf, int64_t* grid, float* depth, void* mask,
                        int64_t* feat_x, int64_t* feat_y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DDF_B200_H_ */
