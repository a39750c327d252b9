"""In-tree build of libddf_b200.so: one nvcc invocation per csrc/*.cu, then a shared link.

sm_100a only (``-gencode arch=compute_100a,code=sm_100a``); nvcc cross-compiles without a GPU.
The built library sits next to this file so it travels with the repo snapshot to the GPU box.
"""
import concurrent.futures as _fut
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libddf_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    m = 0.0
    for d in (CSRC, os.path.join(os.path.dirname(PKG_DIR), "include")):
        for f in os.listdir(d):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def _compile(src, force):
    obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
    spath = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) > max(os.path.getmtime(spath), _headers_mtime())):
        return obj, ""
    cmd = [NVCC, *NVCC_FLAGS, *os.environ.get("DDF_NVCC_EXTRA", "").split(), "-c", spath, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link libddf_b200.so. Returns the library path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = _sources()
    with _fut.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    if (force or not os.path.exists(LIB_PATH)
            or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs)):
        cmd = [NVCC, "-shared", "-o", LIB_PATH, *objs, "-gencode",
               "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
