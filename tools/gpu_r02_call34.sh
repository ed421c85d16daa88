#!/bin/bash
# Round 2, GPU call 34 (1 GPU): ncu --set full of one column-block launch of the weak-scaled R32 shard on the final
# kernel (tile shape 13, per-warp gather rounds).
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_tma -s 1 -c 1 -f -o gpurun_out/r02c34_prof_block python tools/bench_blocks.py --weak --one-block 8 > gpurun_out/r02c34_prof_block.log 2>&1
tail -3 gpurun_out/r02c34_prof_block.log
