#!/bin/bash
# ncu launch list of one step of bench config $1 -> gpurun_out/launches_$1.{csv,md}
mkdir -p gpurun_out
c=${1:-tf}
timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${2:-4200} -c ${3:-1100} --csv --log-file gpurun_out/launches_$c.csv python bench.py --config $c --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$c.log 2>&1; echo "ncu list rc=$?"
python tools/profile_report.py launches gpurun_out/launches_$c.csv gpurun_out/bench_$c.json > gpurun_out/launches_$c.md
head -75 gpurun_out/launches_$c.md
