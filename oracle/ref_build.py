"""Build the REFERENCE's own extensions, unmodified, from the sources where they lie under
/root/reference into oracle/_ref/ (git-ignored; travels to the GPU box like our own .so files).

ORACLE = test infrastructure: used to validate the C/numpy restatements and as the CPU baseline
("kind": "reference") in bench.py. Never imported by the product package. No reference SOURCE is
copied into the repo: the compiler reads the files in place; only binaries land in oracle/_ref/.

    voxel_layer      TransFusion/mmdet3d/ops/voxel/src/*           (hard/dynamic voxelization, CPU+CUDA)
    sparse_conv_ext  TransFusion/mmdet3d/ops/spconv/src/* + include (vendored spconv v1, CPU+CUDA)

Recipe = torch.utils.cpp_extension.load with the same source lists and flags as the reference's
setup.py:171-207 (-w -std=c++14 replaced by c++17, which torch 2.11 headers require), arch 10.0.

Four more extensions exist only as the GPU "kernel to beat" (tools/bench_reference_kernels.py; they have no CPU
path, so they are never an oracle). They do not compile against torch 2.11 as they are; the
recipe compiles a scratch copy under oracle/_build/ (git-ignored) with the two-line fixes below and nothing else:

    MultiScaleDeformableAttention  TransFusion/mmdet3d/models/model_utils/ops/src/**   `value.type()` ->
                                   `value.scalar_type()` in AT_DISPATCH_FLOATING_TYPES (ms_deform_attn_cuda.cu:64,134)
    furthest_point_sample_ext, ball_query_ext, group_points_ext, gather_points_ext
                                   TransFusion/mmdet3d/ops/<op>/src/*: `#include <THC/THC.h>` -> `#include
                                   <ATen/cuda/CUDAContext.h>` (THC is gone from torch 2.x; the files only need
                                   getCurrentCUDAStream from it), `extern THCState *state;` (unused) removed
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE = "/root/reference"
OPS = os.path.join(REFERENCE, "TransFusion", "mmdet3d", "ops")

EXTS = {
    "voxel_layer": dict(
        sources=["voxel/src/voxelization.cpp", "voxel/src/scatter_points_cpu.cpp",
                 "voxel/src/scatter_points_cuda.cu", "voxel/src/voxelization_cpu.cpp",
                 "voxel/src/voxelization_cuda.cu"],
        include=[]),
    "sparse_conv_ext": dict(
        sources=["spconv/src/all.cc", "spconv/src/reordering.cc", "spconv/src/reordering_cuda.cu",
                 "spconv/src/indice.cc", "spconv/src/indice_cuda.cu", "spconv/src/maxpool.cc",
                 "spconv/src/maxpool_cuda.cu"],
        include=["spconv/include"]),
}


MU_OPS = os.path.join(REFERENCE, "TransFusion", "mmdet3d", "models", "model_utils", "ops", "src")
PATCHED = {
    "MultiScaleDeformableAttention": dict(
        root=MU_OPS, files=["vision.cpp", "ms_deform_attn.h", "cpu/ms_deform_attn_cpu.cpp", "cpu/ms_deform_attn_cpu.h",
                            "cuda/ms_deform_attn_cuda.cu", "cuda/ms_deform_attn_cuda.h", "cuda/ms_deform_im2col_cuda.cuh"],
        sources=["vision.cpp", "cpu/ms_deform_attn_cpu.cpp", "cuda/ms_deform_attn_cuda.cu"],
        subs=[("AT_DISPATCH_FLOATING_TYPES(value.type(),", "AT_DISPATCH_FLOATING_TYPES(value.scalar_type(),")]),
}
for _op in ("furthest_point_sample", "ball_query", "group_points", "gather_points"):
    PATCHED[_op + "_ext"] = dict(
        root=os.path.join(OPS, _op, "src"), files=[_op + ".cpp", _op + "_cuda.cu"], sources=[_op + ".cpp", _op + "_cuda.cu"],
        subs=[("#include <THC/THC.h>", "#include <ATen/cuda/CUDAContext.h>"), ("extern THCState *state;", ""),
              ("extern THCState* state;", "")])


def build_patched(name, verbose=False):
    """Scratch copy + the documented two-line fixes, compiled for sm_100 into oracle/_ref/<name>.so."""
    from torch.utils import cpp_extension
    import shutil
    cfg = PATCHED[name]
    os.makedirs(REF_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "8")
    bdir = os.path.join(HERE, "_build", "ref_" + name)
    src = os.path.join(bdir, "src")
    for f in cfg["files"]:
        os.makedirs(os.path.dirname(os.path.join(src, f)), exist_ok=True)
        text = open(os.path.join(cfg["root"], f)).read()
        for a, b in cfg["subs"]:
            text = text.replace(a, b)
        open(os.path.join(src, f), "w").write(text)
    cpp_extension.load(
        name=name, sources=[os.path.join(src, f) for f in cfg["sources"]], extra_include_paths=[src],
        extra_cflags=["-w", "-std=c++17", "-DWITH_CUDA"],
        extra_cuda_cflags=["-w", "-std=c++17", "-DWITH_CUDA", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                           "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"],
        build_directory=bdir, verbose=verbose, is_python_module=False)
    shutil.copy2(os.path.join(bdir, name + ".so"), so_path(name))
    return so_path(name)


def so_path(name):
    return os.path.join(REF_DIR, name + ".so")


def build_ext(name, verbose=False):
    from torch.utils import cpp_extension
    cfg = EXTS[name]
    os.makedirs(REF_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "8")
    bdir = os.path.join(HERE, "_build", "ref_" + name)
    os.makedirs(bdir, exist_ok=True)
    cpp_extension.load(
        name=name, sources=[os.path.join(OPS, s) for s in cfg["sources"]],
        extra_include_paths=[os.path.join(OPS, i) for i in cfg["include"]],
        extra_cflags=["-w", "-std=c++17", "-DWITH_CUDA"],
        extra_cuda_cflags=["-w", "-std=c++17", "-DWITH_CUDA", "-D__CUDA_NO_HALF_OPERATORS__",
                           "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"],
        build_directory=bdir, verbose=verbose, is_python_module=False)
    import shutil
    shutil.copy2(os.path.join(bdir, name + ".so"), so_path(name))
    return so_path(name)


def build_all(only_if_reference_present=True, verbose=False):
    if not os.path.isdir(REFERENCE):
        if only_if_reference_present:
            return []
        raise RuntimeError("/root/reference not present")
    built = []
    for name in EXTS:
        if not os.path.exists(so_path(name)):
            build_ext(name, verbose)
        built.append(so_path(name))
    for name in PATCHED:    # GPU-only "kernel to beat" builds: optional, a failure must not break build()
        try:
            if not os.path.exists(so_path(name)):
                build_patched(name, verbose)
            built.append(so_path(name))
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("oracle/ref_build: %s not built (%s)\n" % (name, str(e).splitlines()[-1] if str(e) else e))
    return built


def load(name):
    """Import a prebuilt reference extension from oracle/_ref (None if it was never built)."""
    p = so_path(name)
    if not os.path.exists(p):
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


if __name__ == "__main__":
    print(build_all(only_if_reference_present=False, verbose="-v" in sys.argv))
