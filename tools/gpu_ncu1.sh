#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spconv_tma_kernel -s 5 -c 1 -f -o gpurun_out/prof_tma python tools/bench_ops.py spconv --iters 1 --warm 0 --stages subm4 > gpurun_out/ncu_tma.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_tma.ncu-rep --page raw --csv > gpurun_out/prof_tma.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_tma.ncu-rep --page source --csv > gpurun_out/prof_tma.source.csv 2>/dev/null
ls -la gpurun_out/prof_tma*
