#!/bin/bash
mkdir -p gpurun_out
export DDF_LIB_PATH=$PWD/3d-dual-fusion_b200/libddf_b200_phases.so
for s in "32->32" "64->64" "128->128"; do timeout 200 python tools/phase_run.py "$s" 2>&1 | grep -v Warning | tail -12; done | tee gpurun_out/phases.log
