#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/check_tc.py > gpurun_out/check_tc.log 2>&1; echo "check_tc rc=$?"; grep -c "e-04" gpurun_out/check_tc.log
timeout 600 python tools/bench_ops.py spconv --tc-mode 1 --stages subm2,subm3 > gpurun_out/bench_spconv_tma.jsonl 2> gpurun_out/bench_ops.err; echo "bench_ops rc=$?"; tail -3 gpurun_out/bench_ops.err
grep "fwd\|dgrad" gpurun_out/bench_spconv_tma.jsonl | cut -c1-100
