#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --config tf --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_tf.err | tee gpurun_out/bench_tf.json | python tools/print_bench.py
timeout 600 python tools/cpu_profile.py --config tf > gpurun_out/cpu_profile_tf.log 2>&1
grep -A28 "Ordered by: internal time" gpurun_out/cpu_profile_tf.log | head -36 | cut -c1-150
