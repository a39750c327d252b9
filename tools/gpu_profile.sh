#!/bin/bash
# One gpurun call: headline bench (with CPU baseline), ncu launch list of the step, ncu --set full of the
# conv / wgrad / MSDA / BatchNorm kernels. Captures are exported to CSV on the box (64 MiB copy-back limit).
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3500 -c 1300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'spconv_tma_kernel|spconv_wgrad|spconv_tc_kernel' -c 24 -f -o gpurun_out/prof_spconv python tools/bench_ops.py spconv --iters 1 --warm 0 --stages subm1,subm3,subm4 > gpurun_out/ncu_spconv.log 2>&1; echo "ncu spconv rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'msda_' -c 2 -f -o gpurun_out/prof_msda python tools/bench_ops.py msda --iters 1 --warm 0 > gpurun_out/ncu_msda.log 2>&1; echo "ncu msda rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'bn_|vox_|subm_tables|dense_scatter' -c 12 -f -o gpurun_out/prof_misc python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_misc.log 2>&1; echo "ncu misc rc=$?"
for f in prof_spconv prof_msda prof_misc; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
  rm -f gpurun_out/$f.ncu-rep
done
timeout 300 python tools/bench_ops.py msda > gpurun_out/bench_msda.jsonl 2>/dev/null
timeout 300 python tools/bench_ops.py voxel > gpurun_out/bench_misc.jsonl 2>/dev/null
timeout 300 python tools/bench_ops.py dense >> gpurun_out/bench_misc.jsonl 2>/dev/null
du -sh gpurun_out
