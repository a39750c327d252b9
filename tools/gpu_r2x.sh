#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msda_tile_gpu.py tests/test_msda_gpu.py -m gpu -q -x > gpurun_out/pytest_msda.log 2>&1; echo "pytest msda rc=$?"
tail -3 gpurun_out/pytest_msda.log
for p in 0 1; do for sh in ctf ccp kitti; do
DDF_MSDA_PERSISTENT=$p timeout 200 python tools/bench_ops.py msda --shape $sh --iters 20 2>&1 | grep "msda_tile_fwd" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('persistent=$p', '$sh', d['kernel'], round(d['ms_median'], 4), 'frac', round(d['frac'], 3))
"
done; done | tee gpurun_out/msda_persist.log
