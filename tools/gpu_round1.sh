#!/bin/bash
# first GPU shake-out: smoke, full gpu test suite, config sweep, memcheck of the smoke path
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/sweep_spmv.py l5 banded r32 > gpurun_out/sweep_stdout.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweep_stdout.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/memcheck_smoke.log 2>&1; echo "memcheck exit $?" >> gpurun_out/memcheck_smoke.log
tail -5 gpurun_out/smoke.log; tail -30 gpurun_out/pytest_gpu.log; tail -60 gpurun_out/sweep_stdout.log; tail -5 gpurun_out/memcheck_smoke.log
