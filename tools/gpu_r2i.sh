#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
DDF_PROFILE_STACKS=1 timeout 600 python tools/torch_profile.py --config tf > gpurun_out/torch_profile_tf.log 2>&1; echo "profile rc=$?"
