#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sparse_norm_gpu.py -m gpu -q -x > gpurun_out/pytest_bn.log 2>&1; echo "pytest bn rc=$?"
tail -3 gpurun_out/pytest_bn.log
timeout 300 python tools/bench_bn.py 2>&1 | tee gpurun_out/bench_bn.jsonl
