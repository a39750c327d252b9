#!/bin/bash
# full GPU suite + every bench config (single GPU)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
for c in tf tf_cam cp cp_pfatv2 kitti dense200k; do timeout 600 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_$c.err | tee gpurun_out/bench_$c.json | python tools/print_bench.py; done
