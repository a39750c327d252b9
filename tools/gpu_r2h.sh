#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_camera_gpu.py tests/test_wrappers_gpu.py tests/test_wrapper_golden.py tests/test_cp_wrapper_golden.py tests/test_hotpath_gpu.py -m gpu -q > gpurun_out/pytest_new.log 2>&1; echo "pytest new rc=$?"
tail -8 gpurun_out/pytest_new.log
for c in tf tf_cam; do timeout 600 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_$c.err | tee gpurun_out/bench_$c.json | python tools/print_bench.py; done
tail -3 gpurun_out/bench_tf_cam.err
