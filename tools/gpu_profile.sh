#!/bin/bash
# One gpurun call: headline bench (with CPU baseline), ncu launch list of the step, DRAM traffic of the conv
# launches of one step, ncu --set full of the conv / wgrad / MSDA / BatchNorm / fused kernels.
# Captures are exported to CSV on the box (64 MiB copy-back limit).
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1100 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'spconv_tma_kernel|spconv_tc_kernel|spconv_sparse_rows' -c 41 --csv --log-file gpurun_out/conv_traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1; echo "ncu traffic rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'spconv_tma_kernel|spconv_wgrad|spconv_tc_kernel|spconv_sparse_rows' -c 24 -f -o gpurun_out/prof_spconv python tools/bench_ops.py spconv --iters 1 --warm 0 --stages subm1,subm3,subm4 > gpurun_out/ncu_spconv.log 2>&1; echo "ncu spconv rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'msda_' -c 2 -f -o gpurun_out/prof_msda python tools/bench_ops.py msda --iters 1 --warm 0 > gpurun_out/ncu_msda.log 2>&1; echo "ncu msda rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'bn_|vox_|subm_tables|dense_scatter|relu_dropout|add_dropout_ln' -c 16 -f -o gpurun_out/prof_misc python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_misc.log 2>&1; echo "ncu misc rc=$?"
for f in prof_spconv prof_msda prof_misc; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
  rm -f gpurun_out/$f.ncu-rep
done
timeout 300 python tools/bench_ops.py spconv > gpurun_out/bench_spconv.jsonl 2>/dev/null
timeout 300 python tools/bench_ops.py msda > gpurun_out/bench_msda.jsonl 2>/dev/null
du -sh gpurun_out
