#!/bin/bash
# Round 2, GPU call 15 (1 GPU): larger tiles (cfg 14: 8 warps x 8 groups, cfg 15: 16 warps x 4 groups) for scattered
# short rows, fp32 and fp64.
mkdir -p gpurun_out
timeout 300 python tools/bench_blocks.py --weak 8 --cfgs=12,13,14,15 > gpurun_out/r02c15_blocks_weak8.log 2>&1
rm -f gpurun_out/sweep_spmv.txt
SWEEP_CFGS=7,9,8,12,13,14,15 timeout 400 python tools/sweep_spmv.py r4 r4f64 > gpurun_out/r02c15_sweep_r4.log 2>&1; cp gpurun_out/sweep_spmv.txt gpurun_out/r02c15_sweep_r4.txt
cat gpurun_out/r02c15_blocks_weak8.log
grep -v "^#" gpurun_out/r02c15_sweep_r4.txt | cut -c1-160
tail -3 gpurun_out/r02c15_sweep_r4.log
