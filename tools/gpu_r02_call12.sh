#!/bin/bash
# Round 2, GPU call 12 (1 GPU): SpMM window kernel with several 16-byte packs per lane, L2 fetch granularity vs the
# weak-scaled R32 shard, and the column-block kernel under ncu.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_zspmm.py tests/test_gpu_spmv.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r02c12_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c12_pytest.log
for np in 1 2 4; do
  B2S_SPMM_NP=$np timeout 300 python tools/bench_spmm.py 4000000 16 32 64 > gpurun_out/r02c12_bench_spmm_np$np.log 2>&1
  cp gpurun_out/spmm_bench.json gpurun_out/r02c12_spmm_bench_np$np.json 2>/dev/null
done
timeout 300 python tools/bench_spmm.py 4000000 8 16 32 64 128 > gpurun_out/r02c12_bench_spmm.log 2>&1; cp gpurun_out/spmm_bench.json gpurun_out/r02c12_spmm_bench.json 2>/dev/null
timeout 400 python tools/bench_blocks.py --l2gran 8 > gpurun_out/r02c12_l2gran.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_tma -s 1 -c 1 -f -o gpurun_out/r02c12_prof_block python tools/bench_blocks.py --weak --one-block 8 > gpurun_out/r02c12_prof_block.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmm_window_tma -c 1 -f -o gpurun_out/r02c12_prof_spmm_tma python tools/prof_spmm.py 4000000 32 > gpurun_out/r02c12_prof_spmm.log 2>&1
tail -3 gpurun_out/r02c12_pytest.log
for np in 1 2 4; do echo "NP=$np"; grep SPMM gpurun_out/r02c12_bench_spmm_np$np.log | cut -c1-250 | head -8; done
cat gpurun_out/r02c12_l2gran.log | tail -8
