#!/bin/bash
# Round 2, GPU call 16 (1 GPU): host-vector product with y stored straight into pinned host memory.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spmv.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r02c16_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c16_pytest.log
timeout 200 python tools/e2e_direct.py > gpurun_out/r02c16_e2e_direct.log 2>&1
timeout 200 python tools/bench_blocks.py --weak 8 --cfgs=-1 > gpurun_out/r02c16_blocks_weak8.log 2>&1
tail -3 gpurun_out/r02c16_pytest.log; cat gpurun_out/r02c16_e2e_direct.log; cat gpurun_out/r02c16_blocks_weak8.log
