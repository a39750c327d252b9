#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/check_tc.py > gpurun_out/check_tc.log 2>&1; echo "check_tc rc=$?"; tail -18 gpurun_out/check_tc.log | cut -c1-80
timeout 600 python tools/bench_ops.py spconv --tc-mode 1 > gpurun_out/bench_spconv_tma.jsonl 2> gpurun_out/bench_ops.err; echo "bench_ops rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
