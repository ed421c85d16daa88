#!/bin/bash
# Round 2, GPU call 30 (2 GPUs): multi-GPU worker (incl. dist.spgemm) on the final SpGEMM kernels, bench at N=2.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dist.py tests/test_gpu_c_abi_sharded.py -x -q -p no:cacheprovider > gpurun_out/r02c30_pytest_dist.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c30_pytest_dist.log
tail -3 gpurun_out/r02c30_pytest_dist.log
