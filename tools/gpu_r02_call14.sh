#!/bin/bash
# Round 2, GPU call 14 (1 GPU): 8-warp tile shapes with L2-coherent gathers (cfg 12, 13) for scattered short rows,
# SpMM with the unrolled window pre-pass.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_zspmm.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r02c14_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c14_pytest.log
timeout 300 python tools/bench_spmm.py 4000000 16 32 64 > gpurun_out/r02c14_bench_spmm.log 2>&1; cp gpurun_out/spmm_bench.json gpurun_out/r02c14_spmm_bench.json 2>/dev/null
timeout 300 python tools/bench_blocks.py --weak 8 --cfgs=7,8,6,11,12,13 > gpurun_out/r02c14_blocks_weak8.log 2>&1
timeout 300 python tools/bench_blocks.py --weak 4 --cfgs=7,8,6,11,12,13 > gpurun_out/r02c14_blocks_weak4.log 2>&1
timeout 300 python tools/bench_blocks.py --weak 2 --cfgs=7,8,6,11,12,13 > gpurun_out/r02c14_blocks_weak2.log 2>&1
rm -f gpurun_out/sweep_spmv.txt
SWEEP_CFGS=5,7,8,6,11,12,13 timeout 300 python tools/sweep_spmv.py r4 > gpurun_out/r02c14_sweep_r4.log 2>&1; cp gpurun_out/sweep_spmv.txt gpurun_out/r02c14_sweep_r4.txt
tail -2 gpurun_out/r02c14_pytest.log
grep SPMM gpurun_out/r02c14_bench_spmm.log | cut -c1-250 | head -6
cat gpurun_out/r02c14_blocks_weak8.log gpurun_out/r02c14_blocks_weak4.log gpurun_out/r02c14_blocks_weak2.log
grep -v "^#" gpurun_out/r02c14_sweep_r4.txt | cut -c1-160
