#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_2gpu.err | tee gpurun_out/bench_2gpu.json | python tools/print_bench.py
tail -2 gpurun_out/bench_2gpu.err
