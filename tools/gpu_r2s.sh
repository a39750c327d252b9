#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spconv_gpu.py tests/test_sparse_norm_gpu.py -m gpu -q -x > gpurun_out/pytest_conv.log 2>&1; echo "pytest conv+bn rc=$?"
tail -4 gpurun_out/pytest_conv.log
timeout 300 python tools/bench_ops.py spconv --iters 10 --stages "32->32,64->64,128->128" 2>&1 | grep "spconv wgrad" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['kernel'], round(d['ms_median'], 4), round(d['TFLOPs'], 1))
" | tee gpurun_out/spconv_wgrad.log
timeout 600 python bench.py --config tf --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_tf.err | tee gpurun_out/bench_tf.json | python tools/print_bench.py
