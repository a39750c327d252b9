#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/alloc_probe.py --config tf > gpurun_out/alloc_probe.log 2>&1; echo "rc=$?"
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True timeout 600 python tools/alloc_probe.py --config tf > gpurun_out/alloc_probe_exp.log 2>&1; echo "rc=$?"
