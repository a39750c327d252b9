#!/bin/bash
mkdir -p gpurun_out
DDF_LIB_PATH=$PWD/3d-dual-fusion_b200/libddf_b200_trace.so timeout 300 python tools/bench_ops.py spconv --stages "32->32,64->64" --iters 1 --warm 0 > gpurun_out/trace_wgrad.log 2>&1; echo "trace rc=$?"
grep -c "^wgrad" gpurun_out/trace_wgrad.log
