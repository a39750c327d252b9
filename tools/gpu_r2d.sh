#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msda_tile_gpu.py -m gpu -q -x > gpurun_out/pytest_new.log 2>&1; echo "pytest new rc=$?"
tail -5 gpurun_out/pytest_new.log
rm -f gpurun_out/bench_msda.jsonl
for s in ctf ccp kitti; do timeout 300 python tools/bench_ops.py msda --shape $s 2>gpurun_out/bench_msda.err | grep tile >> gpurun_out/bench_msda.jsonl; done
DDF_MSDA_STAGE=tma timeout 300 python tools/bench_ops.py msda --shape ctf 2>>gpurun_out/bench_msda.err | grep tile | sed 's/msda_tile/TMA_msda_tile/' >> gpurun_out/bench_msda.jsonl
cut -c1-200 gpurun_out/bench_msda.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])
print(d['kernels'])
PY
NCU_COUNT=2 bash tools/gpu_ncu_one.sh msda_tile prof_msda_tile python tools/bench_ops.py msda --iters 1 --warm 0 | head -8
