#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1050 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
