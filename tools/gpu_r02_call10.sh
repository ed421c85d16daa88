#!/bin/bash
# Round 2, GPU call 10 (1 GPU): final-code validation -- whole -m gpu suite, SpMM kernels, bench + reference arm, ncu of
# the L5 kernel (short-row tile shape cfg 11) and of the TMA SpMM kernel, launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02c10_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c10_pytest_gpu.log
timeout 300 python tools/bench_spmm.py 4000000 16 32 64 > gpurun_out/r02c10_bench_spmm.log 2>&1; cp gpurun_out/spmm_bench.json gpurun_out/r02c10_spmm_bench.json 2>/dev/null
timeout 600 python bench.py --steps 1000 --warmup 20 > gpurun_out/r02c10_bench_n1.json 2> gpurun_out/r02c10_bench_n1.err
timeout 200 python bench.py --impl reference --steps 50 --warmup 10 > gpurun_out/r02c10_bench_ref.json 2> gpurun_out/r02c10_bench_ref.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmv_tma -s 3 -c 1 -f -o gpurun_out/r02c10_prof_spmv_l5 python tools/prof_spmv.py l5 -1 0 5 > gpurun_out/r02c10_prof_l5.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmm_window_tma -c 1 -f -o gpurun_out/r02c10_prof_spmm_tma python tools/prof_spmm.py 4000000 32 > gpurun_out/r02c10_prof_spmm.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 40 --csv --log-file gpurun_out/r02c10_launches_bench_n1.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/r02c10_launches_stdout.log 2>&1
SWEEP_CFGS=0,11 timeout 100 python tools/sweep_spmv.py l5 banded > gpurun_out/r02c10_sweep.log 2>&1
tail -4 gpurun_out/r02c10_pytest_gpu.log; grep SPMM gpurun_out/r02c10_bench_spmm.log | cut -c1-330 | head -6; head -c 700 gpurun_out/r02c10_bench_n1.json; tail -4 gpurun_out/r02c10_sweep.log
