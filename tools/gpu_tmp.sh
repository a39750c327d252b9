#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 600 python bench.py --config tf --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_tf.err | tee gpurun_out/bench_tf.json | python tools/print_bench.py | cut -c1-110; python -c "
import json; d=json.load(open('gpurun_out/bench_tf.json')); print(d['step_ms_rank0'])"; done
