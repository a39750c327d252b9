#!/bin/bash
# One gpurun call: full GPU test suite, per-layer sizes, op micro-benches, ncu --set full captures.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python tools/layer_sizes.py > gpurun_out/layer_sizes.log 2>&1; echo "layer_sizes rc=$?"
timeout 600 python tools/bench_ops.py all > gpurun_out/bench_ops.jsonl 2> gpurun_out/bench_ops.err; echo "bench_ops rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spconv_tc_kernel|spconv_wgrad_tc_kernel' -c 84 -f -o gpurun_out/prof_spconv python tools/bench_ops.py spconv --iters 1 > gpurun_out/ncu_spconv.log 2>&1; echo "ncu spconv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'msda_' -c 8 -f -o gpurun_out/prof_msda python tools/bench_ops.py msda --iters 1 > gpurun_out/ncu_msda.log 2>&1; echo "ncu msda rc=$?"
ls -la gpurun_out
