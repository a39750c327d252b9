#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -x -q -k "ffn" 2>&1 | tail -15
timeout 300 python tools/bench_ffn.py 2>&1 | tee gpurun_out/bench_ffn.jsonl
timeout 600 python bench.py --config tf --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_tf.err | tee gpurun_out/bench_tf.json | python tools/print_bench.py
