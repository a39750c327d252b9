#!/bin/bash
# round 2 call B: new tests first (fail fast), tile MSDA micro-bench, full GPU suite, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msda_tile_gpu.py tests/test_actr_golden.py tests/test_wrapper_golden.py -m gpu -q -x > gpurun_out/pytest_new.log 2>&1; echo "pytest new rc=$?"
tail -25 gpurun_out/pytest_new.log
for s in ctf ccp kitti; do timeout 300 python tools/bench_ops.py msda --shape $s >> gpurun_out/bench_msda.jsonl 2>gpurun_out/bench_msda.err; done
cat gpurun_out/bench_msda.jsonl | cut -c1-220
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])
print(d['kernels'])
PY
