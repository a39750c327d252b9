#!/bin/bash
# end-of-round record: full GPU suite, smoke, every bench config, ncu launch list of one tf step, ncu --set full of the
# round's new kernels
mkdir -p gpurun_out
bash tools/gpu_suite_all_configs.sh
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_tf.csv python tools/step_only.py --config tf > gpurun_out/bench_under_ncu_tf.log 2>&1; echo "ncu list rc=$?"
python tools/profile_report.py launches gpurun_out/launches_tf.csv gpurun_out/bench_tf.json 0.25 > gpurun_out/launches_tf.md
head -14 gpurun_out/launches_tf.md
NCU_COUNT=6 bash tools/gpu_ncu_one.sh "gn_rows|nchw_to_rows|bn_stats|bn_bwd_reduce" prof_rows python tools/step_only.py --config tf > gpurun_out/prof_rows.out 2>&1
head -40 gpurun_out/prof_rows.md
