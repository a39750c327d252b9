#!/bin/bash
mkdir -p gpurun_out
DDF_STEPS=1 NCU_COUNT=9 bash tools/gpu_ncu_one.sh "gn_rows|nchw_to_rows" prof_rows2 python tools/step_only.py --config tf > gpurun_out/prof_rows2.out 2>&1
head -30 gpurun_out/prof_rows2.md
