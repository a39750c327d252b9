#!/bin/bash
# N=2 data-parallel bench: GradientExchange (default) and, with DDP=1, torch's wrapper beside it
mkdir -p gpurun_out
n=${N:-2}
run() {
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 8 --warmup 3 --no-cpu-baseline $2 2>gpurun_out/bench_${n}gpu$1.err | tee gpurun_out/bench_${n}gpu$1.json | python tools/print_bench.py
python -c "
import json; d=json.load(open('gpurun_out/bench_${n}gpu$1.json')); print('rank ms dev', [round(x,2) for x in d['rank_ms_per_step']['device_resident']]); print('rank ms e2e', [round(x,2) for x in d['rank_ms_per_step']['e2e']])"
}
run "" ""
if [ -n "$DDP" ]; then run _ddp --ddp; fi
