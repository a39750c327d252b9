#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fused_gpu.py tests/test_cp_wrapper_golden.py tests/test_wrappers_gpu.py tests/test_ifat.py tests/test_hotpath_gpu.py -m gpu -q -x 2>&1 | tail -3
for c in cp cp_pfatv2; do timeout 600 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_$c.err | tee gpurun_out/bench_$c.json | python tools/print_bench.py | cut -c1-110; done
