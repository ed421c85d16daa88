#!/bin/bash
# Round 2, GPU call 36 (1 GPU): the whole -m gpu suite on the last build of the round.
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02c36_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c36_pytest_gpu.log
tail -3 gpurun_out/r02c36_pytest_gpu.log
