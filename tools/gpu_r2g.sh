#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pointops_gpu.py tests/test_actr_golden.py tests/test_wrappers_gpu.py -m gpu -q -x > gpurun_out/pytest_new.log 2>&1; echo "pytest new rc=$?"
tail -12 gpurun_out/pytest_new.log
for c in kitti cp_pfatv2; do
  timeout 900 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$c.json'))
    print('$c', round(d['value'],2), 'samples/s', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['value'],2), 'launches', d['gpu_launches'])
except Exception as e:
    print('$c failed', e)
PY
done
tail -3 gpurun_out/bench_kitti.err
