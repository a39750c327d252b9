#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spconv_tma_kernel' -c 8 -f -o gpurun_out/prof_conv python tools/bench_ops.py spconv --iters 1 --warm 0 --stages subm2,subm3,subm4 > gpurun_out/prof_conv.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_conv.ncu-rep --page raw --csv > gpurun_out/prof_conv.raw.csv 2>/dev/null
python tools/profile_report.py kernels gpurun_out/prof_conv.raw.csv > gpurun_out/prof_conv.md
cat gpurun_out/prof_conv.md
