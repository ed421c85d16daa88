#!/bin/bash
# First GPU call of round 2 (1 GPU, ~3 min): everything that was only dry-run on the CPU at the end of round 1.
#   gpurun --timeout 420 -- 'bash tools/gpu_round2_first.sh'
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log
timeout 90 python tools/bench_krylov.py 2048 200 > gpurun_out/r02_bench_krylov.log 2>&1; echo "krylov exit $?" >> gpurun_out/r02_bench_krylov.log
timeout 60 python tools/bench_spmm.py 4000000 8 32 128 > gpurun_out/r02_bench_spmm.log 2>&1
tail -8 gpurun_out/r02_pytest_gpu.log; tail -8 gpurun_out/r02_bench_krylov.log; tail -13 gpurun_out/r02_bench_spmm.log
