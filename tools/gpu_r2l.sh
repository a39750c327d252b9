#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_ops.py spconv --iters 10 2>&1 | grep "spconv " | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['kernel'], round(d['ms_median'], 4), round(d['TFLOPs'], 1))
" | tee gpurun_out/spconv_ops.log
timeout 900 python -m pytest tests/test_spconv_gpu.py -m gpu -q -x > gpurun_out/pytest_conv.log 2>&1; echo "pytest conv rc=$?"
tail -3 gpurun_out/pytest_conv.log
