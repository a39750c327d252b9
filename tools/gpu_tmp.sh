#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for c in tf cp_pfatv2 cp kitti tf_cam; do timeout 600 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_$c.err | tee gpurun_out/bench_$c.json | python tools/print_bench.py | cut -c1-110; done
