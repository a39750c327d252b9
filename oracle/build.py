"""Build the oracle's C restatements into oracle/_build/liboracle.so (gcc, OpenMP).

ORACLE = test infrastructure. Nothing under oracle/ is imported by the product package.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(BUILD_DIR, "liboracle.so")


def _sources():
    return sorted(os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".c"))


def build(force=False):
    os.makedirs(BUILD_DIR, exist_ok=True)
    srcs = _sources()
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) > os.path.getmtime(s) for s in srcs)):
        return LIB_PATH
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           "-std=c11", "-o", LIB_PATH, *srcs, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True))
