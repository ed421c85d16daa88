#!/bin/bash
# Round 2, GPU call 21 (1 GPU): ncu of the SpGEMM dense-row kernel (symbolic + numeric launch) on R-MAT scale 18.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spgemm_dense_kernel -c 2 -f -o gpurun_out/r02c21_prof_spgemm_dense python tools/bench_spgemm.py rmat18 > gpurun_out/r02c21_prof_spgemm.log 2>&1
tail -5 gpurun_out/r02c21_prof_spgemm.log
