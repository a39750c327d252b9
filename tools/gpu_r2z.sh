#!/bin/bash
mkdir -p gpurun_out
for c in kitti cp; do timeout 600 python tools/torch_profile.py --config $c > gpurun_out/torch_profile_$c.log 2>&1; echo "profile $c rc=$?"; done
