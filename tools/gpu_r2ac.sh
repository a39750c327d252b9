#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -m gpu -q -x -k "ffn" > gpurun_out/pytest_ffn.log 2>&1; echo "pytest ffn rc=$?"
tail -25 gpurun_out/pytest_ffn.log
timeout 300 python tools/bench_ffn.py 2>&1 | tail -3 | tee gpurun_out/bench_ffn.jsonl
