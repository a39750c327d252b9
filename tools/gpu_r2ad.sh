#!/bin/bash
mkdir -p gpurun_out
for d in 0 1; do echo "dbg=$d"; DDF_FFN_DBG=$d timeout 300 python tools/bench_ffn.py 2>&1 | grep "fused ffn"; done
NCU_COUNT=2 bash tools/gpu_ncu_one.sh "ffn_fwd_kernel" prof_ffn python tools/bench_ffn.py > /dev/null 2>&1
cat gpurun_out/prof_ffn.md
ncu -i gpurun_out/prof_ffn.ncu-rep --page details 2>/dev/null | grep -E "Stall|Eligible|Issued Warp|DRAM Throughput|L2 Cache Throughput|Mem Busy|Max Bandwidth|Tensor|sectors" | head -30
