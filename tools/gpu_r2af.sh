#!/bin/bash
DDF_LIB_PATH=$PWD/3d-dual-fusion_b200/libddf_b200_trace.so timeout 300 python tools/bench_ffn.py 2>&1 | grep "^ffn" | head -16
