#!/bin/bash
# round-2 profiles: launch list of one tf step + ncu --set full of the main kernels
mkdir -p gpurun_out
timeout 300 python bench.py --config tf --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tf.json 2> gpurun_out/bench_tf.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_tf.csv python tools/step_only.py --config tf > gpurun_out/bench_under_ncu_tf.log 2>&1; echo "ncu list rc=$?"
python tools/profile_report.py launches gpurun_out/launches_tf.csv gpurun_out/bench_tf.json 0.25 > gpurun_out/launches_tf.md
head -12 gpurun_out/launches_tf.md
NCU_COUNT=6 bash tools/gpu_ncu_one.sh "spconv_tma_kernel|spconv_wgrad_table" prof_conv2 python tools/bench_ops.py spconv --stages "32->32,64->64,128->128" --iters 1 --warm 0 > /dev/null 2>&1
NCU_COUNT=4 bash tools/gpu_ncu_one.sh "xty_kernel" prof_xty python tools/bench_xty.py > /dev/null 2>&1
NCU_COUNT=4 bash tools/gpu_ncu_one.sh "msda_tile" prof_msda2 python tools/bench_ops.py msda --shape ctf --iters 1 --warm 0 > /dev/null 2>&1
cat gpurun_out/prof_conv2.md gpurun_out/prof_xty.md gpurun_out/prof_msda2.md
