#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reference_cuda_gpu.py tests/test_pointops_gpu.py -m gpu -q -x > gpurun_out/pytest_new.log 2>&1; echo "pytest new rc=$?"
tail -8 gpurun_out/pytest_new.log
timeout 300 python tools/bench_reference_kernels.py --only pointops > gpurun_out/ref_kernels_pointops.jsonl 2>&1; cut -c1-250 gpurun_out/ref_kernels_pointops.jsonl
for c in kitti cp_pfatv2; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2400 --csv --log-file gpurun_out/launches_$c.csv python bench.py --config $c --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$c.log 2>&1; echo "ncu list $c rc=$?"
python tools/profile_report.py launches gpurun_out/launches_$c.csv > gpurun_out/launches_$c.md 2>/dev/null
head -45 gpurun_out/launches_$c.md
done
