"""Instrumented build of the CUDA library (-DDDF_TRACE: in-kernel clock64 timelines of one CTA of the conv kernels,
schedule overrides through DDF_TMA_T / DDF_TMA_SLOTS / DDF_TMA_SB) next to the product library:
    python tools/build_trace.py && DDF_LIB_PATH=3d-dual-fusion_b200/libddf_b200_trace.so python tools/bench_ops.py spconv ...
Never used by tests or bench.py."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "3d-dual-fusion_b200")
OBJ = os.path.join(PKG, "build", "trace")
os.makedirs(OBJ, exist_ok=True)
flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"] + (sys.argv[1:] or ["-DDDF_TRACE"])
NAME = "tune" if "-DDDF_TUNE" in sys.argv else "phases" if "-DDDF_PHASES" in sys.argv else "trace"
OBJ = os.path.join(PKG, "build", NAME)
os.makedirs(OBJ, exist_ok=True)
objs = []
procs = []
for f in sorted(os.listdir(os.path.join(PKG, "csrc"))):
    if f.endswith(".cu"):
        o = os.path.join(OBJ, f[:-3] + ".o")
        objs.append(o)
        procs.append(subprocess.Popen(["/usr/local/cuda/bin/nvcc", *flags, "-c", os.path.join(PKG, "csrc", f), "-o", o]))
assert all(p.wait() == 0 for p in procs)
out = os.path.join(PKG, "libddf_b200_%s.so" % NAME)
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
print(out)
