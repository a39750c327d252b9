#!/bin/bash
# One gpurun call: GPU test suite, per-layer sizes, op micro-benches, small ncu --set full captures.
# gpurun copies back at most 64 MiB: captures are exported to CSV on the box and big reports dropped.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python tools/layer_sizes.py > gpurun_out/layer_sizes.log 2>&1; echo "layer_sizes rc=$?"
timeout 600 python tools/bench_ops.py all > gpurun_out/bench_ops.jsonl 2> gpurun_out/bench_ops.err; echo "bench_ops rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spconv_tc_kernel|spconv_wgrad_tc_kernel' -c 12 -f -o gpurun_out/prof_spconv python tools/bench_ops.py spconv --iters 1 --warm 0 --stages subm3,subm4 > gpurun_out/ncu_spconv.log 2>&1; echo "ncu spconv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'msda_' -c 2 -f -o gpurun_out/prof_msda python tools/bench_ops.py msda --iters 1 --warm 0 > gpurun_out/ncu_msda.log 2>&1; echo "ncu msda rc=$?"
for f in prof_spconv prof_msda; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/$f.ncu-rep); if [ "$sz" -gt 25000000 ]; then rm gpurun_out/$f.ncu-rep; echo "dropped $f.ncu-rep ($sz bytes)"; fi
done
du -sh gpurun_out; ls -la gpurun_out
