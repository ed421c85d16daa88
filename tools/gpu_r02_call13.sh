#!/bin/bash
# Round 2, GPU call 13 (1 GPU): short-row path with per-warp round counts (ragged column blocks), SpMM window kernel
# with the tile windows found by the producer warp, block-count / tile-shape sweep for the weak-scaled R32 shard.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_zspmm.py tests/test_gpu_spmv.py tests/test_gpu_cg.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r02c13_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c13_pytest.log
timeout 300 python tools/bench_spmm.py 4000000 16 32 64 > gpurun_out/r02c13_bench_spmm.log 2>&1; cp gpurun_out/spmm_bench.json gpurun_out/r02c13_spmm_bench.json 2>/dev/null
timeout 300 python tools/bench_blocks.py --weak 8 --cfgs=-1,8,4,3,0,10,11,6 > gpurun_out/r02c13_blocks_weak8.log 2>&1
timeout 300 python tools/bench_blocks.py --weak 8 --nblocks=4 --cfgs=-1,8 > gpurun_out/r02c13_blocks_weak8_q4.log 2>&1
timeout 300 python tools/bench_blocks.py --weak 8 --nblocks=16 --cfgs=-1,8 > gpurun_out/r02c13_blocks_weak8_q16.log 2>&1
timeout 300 python tools/bench_blocks.py --weak 4 --cfgs=-1,8 > gpurun_out/r02c13_blocks_weak4.log 2>&1
timeout 300 python tools/bench_blocks.py 2 4 8 > gpurun_out/r02c13_blocks_strong.log 2>&1
B2S_BENCH_EXTRAS=r32,spmm,cg timeout 400 python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/r02c13_bench_n1.json 2> gpurun_out/r02c13_bench_n1.err
tail -3 gpurun_out/r02c13_pytest.log
grep SPMM gpurun_out/r02c13_bench_spmm.log | cut -c1-250 | head -6
cat gpurun_out/r02c13_blocks_weak8.log gpurun_out/r02c13_blocks_weak8_q4.log gpurun_out/r02c13_blocks_weak8_q16.log gpurun_out/r02c13_blocks_weak4.log
head -c 700 gpurun_out/r02c13_bench_n1.json
