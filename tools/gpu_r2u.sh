#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_hotpath_gpu.py tests/test_actr_golden.py -m gpu -q -x > gpurun_out/pytest_enc.log 2>&1; echo "pytest enc rc=$?"
tail -3 gpurun_out/pytest_enc.log
timeout 600 python tools/torch_profile.py --config tf > gpurun_out/torch_profile_tf3.log 2>&1; echo "profile rc=$?"
