#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -m gpu -q -x > gpurun_out/pytest_fused.log 2>&1; echo "pytest fused rc=$?"
tail -15 gpurun_out/pytest_fused.log
timeout 300 python tools/bench_xty.py 2>&1 | tee gpurun_out/bench_xty.jsonl
