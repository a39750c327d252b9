#!/bin/bash
# round 2 call C: tile MSDA after prologue pipelining: tests, micro-bench, ncu --set full of old + tile kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msda_tile_gpu.py tests/test_msda_gpu.py -m gpu -q -x > gpurun_out/pytest_new.log 2>&1; echo "pytest new rc=$?"
tail -5 gpurun_out/pytest_new.log
rm -f gpurun_out/bench_msda.jsonl
for s in ctf ccp kitti; do timeout 300 python tools/bench_ops.py msda --shape $s >> gpurun_out/bench_msda.jsonl 2>gpurun_out/bench_msda.err; done
cut -c1-200 gpurun_out/bench_msda.jsonl
timeout 600 ncu --set full --clock-control none -k regex:'msda_' -c 8 -f -o gpurun_out/prof_msda python tools/bench_ops.py msda --iters 1 --warm 0 > gpurun_out/ncu_msda.log 2>&1; echo "ncu msda rc=$?"
ncu -i gpurun_out/prof_msda.ncu-rep --page raw --csv > gpurun_out/prof_msda.raw.csv 2>/dev/null
python tools/profile_report.py kernels gpurun_out/prof_msda.raw.csv > gpurun_out/prof_msda.md
cat gpurun_out/prof_msda.md
rm -f gpurun_out/prof_msda.ncu-rep
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])
print(d['kernels'])
PY
