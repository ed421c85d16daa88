#!/bin/bash
# Round 2, GPU call 7 (1 GPU): whole -m gpu suite on the final code, bench + reference arm, SpMM window kernel (bench +
# ncu), graded e2e pipeline trace, column-block experiment.   gpurun --timeout 1500 -- 'bash tools/gpu_r02_call7.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02c7_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c7_pytest_gpu.log
timeout 200 python tools/bench_spmm.py 4000000 8 32 128 > gpurun_out/r02c7_bench_spmm.log 2>&1; cp gpurun_out/spmm_bench.json gpurun_out/r02c7_spmm_bench.json 2>/dev/null
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmm_tile -s 1 -c 1 -f -o gpurun_out/r02c7_prof_spmm_window python tools/prof_spmm.py 4000000 32 > gpurun_out/r02c7_prof_spmm.log 2>&1
B2S_PIPE_TRACE=1 timeout 100 python tools/e2e_trace.py > gpurun_out/r02c7_e2e_trace.log 2>&1
timeout 300 python tools/bench_blocks.py 8 > gpurun_out/r02c7_bench_blocks.log 2>&1
timeout 600 python bench.py --steps 1000 --warmup 20 > gpurun_out/r02c7_bench_n1.json 2> gpurun_out/r02c7_bench_n1.err
timeout 200 python bench.py --impl reference --steps 50 --warmup 10 > gpurun_out/r02c7_bench_ref.json 2> gpurun_out/r02c7_bench_ref.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 40 --csv --log-file gpurun_out/r02c7_launches_bench_n1.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/r02c7_launches_stdout.log 2>&1
tail -4 gpurun_out/r02c7_pytest_gpu.log; grep SPMM gpurun_out/r02c7_bench_spmm.log | cut -c1-260 | head -6; tail -8 gpurun_out/r02c7_e2e_trace.log; head -8 gpurun_out/r02c7_bench_blocks.log; head -c 900 gpurun_out/r02c7_bench_n1.json
