#!/bin/bash
# Round 2, GPU call 33 (1 GPU): ncu --set full of the final L5 fp64 SpMV kernel (refreshes profiles/spmv_traffic.json).
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_tma -s 3 -c 1 -f -o gpurun_out/r02c33_prof_spmv_l5 python tools/prof_spmv.py l5 -1 0 5 > gpurun_out/r02c33_prof_l5.log 2>&1
tail -3 gpurun_out/r02c33_prof_l5.log
