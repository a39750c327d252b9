#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_tf_last.err | tee gpurun_out/bench_tf_last.json | python tools/print_bench.py | cut -c1-120
python -c "
import json; d=json.load(open('gpurun_out/bench_tf_last.json')); print(d['grad_sync'], d['clocks'], d['gpu_launches'])"
