#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 4 6 8 14; do
echo "== dbg=$d"
DDF_CONV_DBG=$d timeout 300 python tools/bench_ops.py spconv --iters 10 --stages "32->32,64->64,128->128" 2>&1 | grep "spconv fwd subm" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['kernel'], round(d['ms_median'], 4), round(d['TFLOPs'], 1))
"
done | tee gpurun_out/spconv_dbg.log
