#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_actr_golden.py tests/test_hotpath_gpu.py tests/test_pointops_gpu.py tests/test_wrappers_gpu.py -m gpu -q > gpurun_out/pytest_enc.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_enc.log
for c in tf cp_pfatv2; do timeout 600 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_$c.err | tee gpurun_out/bench_$c.json | python tools/print_bench.py; done
