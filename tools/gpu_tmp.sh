#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spconv_gpu.py tests/test_hotpath_gpu.py -x -q 2>&1 | tail -2
for i in 1 2; do timeout 600 python bench.py --config tf --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_tf.err | tee gpurun_out/bench_tf.json | python tools/print_bench.py | cut -c1-110; done
DDF_PROFILE_STACKS=1 timeout 600 python tools/torch_profile.py --config tf > gpurun_out/torch_profile_tf_stacks.log 2>&1; echo rc=$?
