#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/check_tc.py > gpurun_out/check_tc.log 2>&1; echo "check_tc rc=$?"; grep -v Warn gpurun_out/check_tc.log | tail -18 | cut -c1-90
timeout 600 python tools/bench_ops.py spconv --tc-mode 1 --stages subm1 > gpurun_out/bench_spconv_tma.jsonl 2> gpurun_out/bench_ops.err; echo "bench_ops rc=$?"; tail -3 gpurun_out/bench_ops.err
grep "fwd\|dgrad" gpurun_out/bench_spconv_tma.jsonl | cut -c1-100
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-330 gpurun_out/bench.json
