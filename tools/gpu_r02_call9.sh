#!/bin/bash
# Round 2, GPU call 9 (1 GPU): SpMM with the persistent TMA X-window kernel (tests + timings + ncu).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zspmm.py tests/test_gpu_spgemm.py -x -q -p no:cacheprovider > gpurun_out/r02c9_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c9_pytest.log
timeout 300 python tools/bench_spmm.py 4000000 8 32 64 > gpurun_out/r02c9_bench_spmm.log 2>&1; cp gpurun_out/spmm_bench.json gpurun_out/r02c9_spmm_bench.json 2>/dev/null
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmm_window_tma -c 1 -f -o gpurun_out/r02c9_prof_spmm_tma python tools/prof_spmm.py 4000000 32 > gpurun_out/r02c9_prof_spmm.log 2>&1
B2S_BENCH_EXTRAS=spmm timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/r02c9_bench_spmm_extra.json 2> gpurun_out/r02c9_bench_spmm_extra.err
tail -15 gpurun_out/r02c9_pytest.log; grep SPMM gpurun_out/r02c9_bench_spmm.log | cut -c1-330; tail -3 gpurun_out/r02c9_prof_spmm.log
