#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -m gpu -q -x -k "ffn" 2>&1 | tail -3
timeout 300 python tools/bench_ffn.py 2>&1 | grep "kernel"
DDF_LIB_PATH=$PWD/3d-dual-fusion_b200/libddf_b200_trace.so timeout 300 python tools/bench_ffn.py 2>&1 | grep "^ffn" | head -16 | tee gpurun_out/trace_ffn.log
