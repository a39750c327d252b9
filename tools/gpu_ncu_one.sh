#!/bin/bash
# ncu --set full of kernels matching $1 (regex) in the command "$2..." ; summary table to gpurun_out/$3.md
mkdir -p gpurun_out
re=$1; out=$2; shift 2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$re" -c ${NCU_COUNT:-4} -f -o gpurun_out/$out "$@" > gpurun_out/$out.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/$out.ncu-rep --page raw --csv > gpurun_out/$out.raw.csv 2>/dev/null
python tools/profile_report.py kernels gpurun_out/$out.raw.csv > gpurun_out/$out.md
cat gpurun_out/$out.md
ncu -i gpurun_out/$out.ncu-rep --page details --csv 2>/dev/null | grep -i -E "stall|Warp Cycles Per|No Eligible|Issue Slot|Theoretical Occ|Achieved Occ|L1/TEX Hit|Mem Busy|Max Bandwidth|Mem Pipes" | cut -c1-260 | head -60
