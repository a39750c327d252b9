#!/bin/bash
mkdir -p gpurun_out
DDF_LIB_PATH=$PWD/3d-dual-fusion_b200/libddf_b200_trace.so timeout 300 python tools/bench_ops.py spconv --stages "32->32,64->64" --iters 1 --warm 0 > gpurun_out/trace32.log 2>&1; echo "trace rc=$?"
export DDF_LIB_PATH=$PWD/3d-dual-fusion_b200/libddf_b200_tune.so
rm -f gpurun_out/sweep.log
for cfg in "2 4 2" "2 4 4" "2 4 5" "2 4 3" "4 8 4" "2 8 4" "1 8 4" "1 4 2"; do
  set -- $cfg
  echo "== T=$1 slots=$2 sb=$3" >> gpurun_out/sweep.log
  DDF_TMA_T=$1 DDF_TMA_SLOTS=$2 DDF_TMA_SB=$3 timeout 300 python tools/bench_ops.py spconv --stages "32->32,64->64" --iters 10 2>&1 | grep "spconv fwd\|spconv dgrad" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['kernel'], round(d['ms_median'], 4), round(d['TFLOPs'], 1))
" >> gpurun_out/sweep.log
done
cat gpurun_out/sweep.log
