#!/bin/bash
# round 2 call A: GPU tests, parity calibration per precision mode, bench in bf16x3 (default) and tf32 mode
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/measure_parity.py 6000 > gpurun_out/parity.jsonl 2> gpurun_out/parity.err; echo "parity rc=$?"
cat gpurun_out/parity.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mode4.json 2> gpurun_out/bench_mode4.err; echo "bench4 rc=$?"
cut -c1-400 gpurun_out/bench_mode4.json
DDF_TC_MODE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mode1.json 2> gpurun_out/bench_mode1.err; echo "bench1 rc=$?"
cut -c1-400 gpurun_out/bench_mode1.json
timeout 300 python tools/bench_ops.py spconv > gpurun_out/bench_spconv_mode4.jsonl 2>/dev/null
timeout 300 python tools/bench_ops.py spconv --tc-mode 1 > gpurun_out/bench_spconv_mode1.jsonl 2>/dev/null
