#!/bin/bash
# round 2 call E: full GPU suite, reference-kernel comparison, bench of every config
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_reference_kernels.py > gpurun_out/ref_kernels.jsonl 2> gpurun_out/ref_kernels.err; echo "refk rc=$?"
cut -c1-330 gpurun_out/ref_kernels.jsonl
for c in tf cp cp_pfatv2 kitti dense200k; do
  timeout 900 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$c.json'))
    print('$c', round(d['value'],2), 'samples/s', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['value'],2), 'launches', d['gpu_launches'], 'conv TF', round(d['roofline']['achieved'],1), {k:(round(v['ms_per_step'],3), round(v.get('frac') or 0,3)) for k,v in d['kernels'].items()})
except Exception as e:
    print('$c failed', e)
PY
done
tail -5 gpurun_out/bench_cp.err
