"""B200-native (sm_100a) implementation of the 3D-Dual-Fusion hot path.

Layout
  csrc/      hand-written CUDA kernels + the C-ABI (``libddf_b200.so``, see include/ddf_b200.h)
  lib.py     ctypes loader of the C-ABI; fails loudly if the library is missing
  build.py   in-tree nvcc build of the library
  ops/       host-side mirrors of the reference's op packages (same names / argument meaning)
  fusion/    the 3D-DF fusion encoder modules with the reference's state-dict layout
  data_parallel.py  gradient averaging over the ranks: one flat NCCL all-reduce per step
"""
__version__ = "0.1.0"
