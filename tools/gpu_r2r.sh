#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/torch_profile.py --config tf > gpurun_out/torch_profile_tf2.log 2>&1; echo "profile rc=$?"
grep "bn_\|Self CUDA time total\|split_bf16\|xty\|cutlass" gpurun_out/torch_profile_tf2.log | cut -c1-60,150-230
