"""ctypes loader for the C-ABI library ``libddf_b200.so`` (see include/ddf_b200.h).

There is NO CPU fallback: if the library is missing or a call fails, a RuntimeError is raised,
mirroring the reference's AT_ASSERTM / TORCH_CHECK behaviour at the pybind boundary.
"""
import ctypes
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# DDF_LIB_PATH: an instrumented build of the same library (tools/build_trace.py), for kernel timeline traces only
LIB_PATH = os.environ.get("DDF_LIB_PATH") or os.path.join(_PKG_DIR, "libddf_b200.so")

_lib = None

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_f32 = ctypes.c_float
c_ptr = ctypes.c_void_p

# name -> argtypes; restype is int (error code) unless listed in _RESTYPES
_SIGNATURES = {
    "ddf_abi_version": [],
    "ddf_compiled_arch": [],
    "ddf_launch_count": [c_int],
    "ddf_ms_deform_attn_forward": [c_ptr] * 6 + [c_i64] * 8 + [c_int, c_ptr],
    "ddf_ms_deform_attn_backward": [c_ptr] * 9 + [c_i64] * 8 + [c_int, c_ptr],
    "ddf_msda_tile_supported": [c_i64] * 4,
    "ddf_msda_plan_bytes": [c_i64] * 4,
    "ddf_msda_plan": [c_ptr, c_ptr, c_ptr] + [c_i64] * 5 + [c_ptr],
    "ddf_msda_tile_forward": [c_ptr] * 6 + [c_i64] * 6 + [c_ptr],
    "ddf_msda_tile_backward": [c_ptr] * 9 + [c_i64] * 6 + [c_ptr],
    "ddf_hard_voxelize_workspace_bytes": [c_i64] * 3,
    "ddf_hard_voxelize": [c_ptr] * 7 + [c_i64] * 4 + [c_ptr, c_i64, c_ptr],
    "ddf_hard_voxelize_mean": [c_ptr] * 7 + [c_i64] * 5 + [c_ptr, c_i64, c_ptr],
    "ddf_dynamic_voxelize": [c_ptr] * 4 + [c_i64] * 2 + [c_ptr],
    "ddf_indice_pairs_workspace_bytes": [c_i64, c_i64] + [c_ptr] * 6 + [c_int],
    "ddf_subm_indice_pairs": [c_ptr, c_i64, c_i64] + [c_ptr] * 8 + [c_i64, c_ptr],
    "ddf_conv_count_outputs": [c_ptr, c_i64, c_i64] + [c_ptr] * 8 + [c_i64, c_ptr],
    "ddf_conv_indice_pairs": [c_ptr, c_i64, c_i64] + [c_ptr] * 6 + [c_i64] + [c_ptr] * 6 + [c_i64, c_ptr],
    "ddf_indice_conv": [c_ptr] * 4 + [c_i64, c_ptr] + [c_i64] * 4 + [c_int, c_int, c_ptr, c_ptr],
    "ddf_indice_conv_backward": [c_ptr] * 5 + [c_i64, c_ptr, c_ptr] + [c_i64] * 4 + [c_int, c_int, c_ptr, c_ptr, c_ptr],
    "ddf_sparse_conv_forward": [c_ptr] * 6 + [c_i64] * 5 + [c_int, c_ptr],
    "ddf_sparse_conv_tc_mode": [c_i64] * 3,
    "ddf_set_tensor_cores": [c_int],
    "ddf_round_tf32": [c_ptr, c_ptr, c_i64, c_ptr],
    "ddf_sparse_conv_dgrad": [c_ptr] * 5 + [c_i64] * 5 + [c_int, c_ptr],
    "ddf_split_bf16x3": [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_ptr],
    "ddf_sparse_conv_wgrad_table": [c_ptr] * 4 + [c_i64] * 5 + [c_ptr],
    "ddf_sparse_conv_wgrad": [c_ptr] * 4 + [c_i64, c_ptr] + [c_i64] * 3 + [c_int, c_ptr],
    "ddf_furthest_point_sampling": [c_ptr] * 3 + [c_i64] * 3 + [c_ptr],
    "ddf_ball_query": [c_ptr] * 3 + [c_i64] * 3 + [c_f32, c_f32, c_i64, c_ptr],
    "ddf_group_points": [c_ptr] * 3 + [c_i64] * 5 + [c_ptr],
    "ddf_group_points_grad": [c_ptr] * 3 + [c_i64] * 5 + [c_ptr],
    "ddf_gather_points": [c_ptr] * 3 + [c_i64] * 4 + [c_ptr],
    "ddf_gather_points_grad": [c_ptr] * 3 + [c_i64] * 4 + [c_ptr],
    "ddf_local_attn_supported": [c_i64] * 3,
    "ddf_local_attn_forward": [c_ptr, c_ptr] + [c_i64] * 4 + [c_ptr],
    "ddf_local_attn_backward": [c_ptr] * 3 + [c_i64] * 4 + [c_ptr],
    "ddf_first_occurrence": [c_ptr, c_ptr] + [c_i64] * 3 + [c_ptr],
    "ddf_scatter_first": [c_ptr] * 3 + [c_i64] * 4 + [c_ptr],
    "ddf_scatter_first_grad": [c_ptr] * 4 + [c_i64] * 4 + [c_ptr],
    "ddf_sparse_to_bev_nhwc_bf16": [c_ptr] * 3 + [c_i64] * 6 + [c_ptr],
    "ddf_bev_nhwc_bf16_to_sparse": [c_ptr] * 3 + [c_i64] * 6 + [c_ptr],
    "ddf_sparse_bn_workspace_bytes": [c_i64],
    "ddf_sparse_bn_forward": [c_ptr] * 9 + [c_i64, c_i64, c_int, c_f32, c_f32, c_int, c_ptr, c_ptr],
    "ddf_sparse_bn_forward_split": [c_ptr] * 11 + [c_i64, c_i64, c_int, c_f32, c_f32, c_int, c_ptr, c_ptr],
    "ddf_sparse_bn_backward": [c_ptr] * 10 + [c_i64, c_i64, c_int, c_int, c_ptr, c_ptr],
    "ddf_bias_relu_dropout_forward": [c_ptr] * 3 + [c_i64, c_i64, c_f32, ctypes.c_uint64, c_ptr],
    "ddf_bias_relu_dropout_backward": [c_ptr] * 4 + [c_i64, c_i64, c_f32, c_ptr],
    "ddf_add_dropout_layer_norm_forward": [c_ptr] * 8 + [c_i64, c_i64, c_f32, ctypes.c_uint64, c_f32, c_ptr],
    "ddf_add_dropout_layer_norm_backward": [c_ptr] * 9 + [c_i64, c_i64, c_f32, ctypes.c_uint64, c_ptr],
    "ddf_project_assign": [c_ptr, c_i64, c_i64, c_ptr, c_i64] + [c_f32] * 9 + [c_i64, c_ptr, c_ptr, c_ptr, c_ptr],
    "ddf_group_ranks": [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr],
    "ddf_project_cameras": [c_ptr] * 6 + [c_i64] * 3 + [c_f32, c_i64, c_i64] + [c_ptr] * 6,
    "ddf_col_sum": [c_ptr, c_ptr, c_i64, c_i64, c_ptr],
    "ddf_bigate_sum_forward": [c_ptr] * 9 + [c_i64, c_i64, c_int, c_ptr],
    "ddf_bigate_sum_backward": [c_ptr] * 13 + [c_i64, c_i64, c_int, c_ptr],
    "ddf_ffn_supported": [c_i64] * 3,
    "ddf_ffn_workspace_bytes": [c_i64] * 2,
    "ddf_ffn_forward": [c_ptr] * 8 + [c_i64] * 3 + [c_f32, ctypes.c_uint64, c_ptr],
    "ddf_ffn_dropout_p": [c_f32],
    "ddf_group_norm_rows_supported": [c_i64] * 2,
    "ddf_nchw_to_rows": [c_ptr, c_int, c_ptr] + [c_i64] * 3 + [c_ptr],
    "ddf_group_norm_rows_forward": [c_ptr] * 7 + [c_i64] * 4 + [c_f32, c_ptr],
    "ddf_group_norm_rows_backward": [c_ptr] * 9 + [c_i64] * 4 + [c_ptr],
    "ddf_xty_supported": [c_i64] * 3,
    "ddf_xty_tf32": [c_ptr] * 3 + [c_i64] * 3 + [c_ptr],
    "ddf_sparse_to_dense": [c_ptr] * 3 + [c_i64] * 6 + [c_ptr],
    "ddf_dense_to_sparse": [c_ptr] * 3 + [c_i64] * 6 + [c_ptr],
}
_RESTYPES = {"ddf_ffn_workspace_bytes": c_i64, "ddf_ffn_dropout_p": c_f32, "ddf_launch_count": c_i64, "ddf_msda_plan_bytes": c_i64, "ddf_hard_voxelize_workspace_bytes": c_i64, "ddf_indice_pairs_workspace_bytes": c_i64,
             "ddf_sparse_bn_workspace_bytes": c_i64}


def exported_symbols():
    """Every symbol include/ddf_b200.h declares (tests check the library exports them all)."""
    return ["ddf_last_error"] + sorted(_SIGNATURES)


def get_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "ddf_b200: %s not found - build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'`. There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.ddf_last_error.restype = ctypes.c_char_p
        lib.ddf_last_error.argtypes = []
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = get_lib().ddf_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream():
    """Raw handle of torch's current stream on the current device. The ``torch.cuda.current_stream()`` object path costs
    about 14 us per call (cProfile of a bench step: 220 calls = 3.8 ms of a 34 ms host-side step); the raw query is
    a fraction of a microsecond."""
    import torch
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


class _NoGuard(object):
    __slots__ = ()

    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def on_device(device):
    """``with on_device(t.device):`` = ``with torch.cuda.device(t.device):`` when the tensor lives on another device than
    the current one, and nothing at all (no context-manager churn) in the usual single-device-per-process case."""
    import torch
    idx = device.index
    if idx is None or idx == torch._C._cuda_getDevice():
        return _NO_GUARD
    return torch.cuda.device(device)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("Not implemented on the CPU")  # reference: ms_deform_attn.h:38
