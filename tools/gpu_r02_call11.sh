#!/bin/bash
# Round 2, GPU call 11 (1 GPU): validation of the final code (suite, bench, reference arm), SpMM timings, and the
# weak-scaling column-block kernel under ncu.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02c11_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c11_pytest_gpu.log
timeout 300 python tools/bench_spmm.py 4000000 16 32 64 > gpurun_out/r02c11_bench_spmm.log 2>&1; cp gpurun_out/spmm_bench.json gpurun_out/r02c11_spmm_bench.json 2>/dev/null
timeout 600 python bench.py --steps 1000 --warmup 20 > gpurun_out/r02c11_bench_n1.json 2> gpurun_out/r02c11_bench_n1.err
timeout 200 python bench.py --impl reference --steps 50 --warmup 10 > gpurun_out/r02c11_bench_ref.json 2> gpurun_out/r02c11_bench_ref.err
timeout 300 python tools/bench_blocks.py --weak 8 > gpurun_out/r02c11_bench_blocks_weak.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_tma -s 1 -c 1 -f -o gpurun_out/r02c11_prof_block python tools/bench_blocks.py --weak --one-block 8 > gpurun_out/r02c11_prof_block.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmm_window_tma -c 1 -f -o gpurun_out/r02c11_prof_spmm_tma python tools/prof_spmm.py 4000000 32 > gpurun_out/r02c11_prof_spmm.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 40 --csv --log-file gpurun_out/r02c11_launches_bench_n1.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/r02c11_launches_stdout.log 2>&1
tail -4 gpurun_out/r02c11_pytest_gpu.log; grep SPMM gpurun_out/r02c11_bench_spmm.log | cut -c1-330 | head -6; head -c 600 gpurun_out/r02c11_bench_n1.json; cat gpurun_out/r02c11_bench_blocks_weak.log
