#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_wrappers_gpu.py tests/test_cp_wrapper_golden.py tests/test_wrapper_golden.py -m gpu -q > gpurun_out/pytest_wrap.log 2>&1; echo "pytest wrappers rc=$?"
tail -6 gpurun_out/pytest_wrap.log
for c in cp cp_pfatv2; do timeout 600 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_$c.err | tee gpurun_out/bench_$c.json | python tools/print_bench.py; done
