#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_msda_tile_gpu.py -x -q 2>&1 | tail -2
for i in 1 2 3; do timeout 600 python bench.py --config tf --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_tf.err | tee gpurun_out/bench_tf.json | python tools/print_bench.py | cut -c1-110; done
timeout 600 python bench.py --config kitti --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_kitti.err | tee gpurun_out/bench_kitti.json | python tools/print_bench.py | cut -c1-110
