#!/bin/bash
# Round 2, GPU call 35 (1 GPU): SpGEMM tests incl. the dense kernel's global level-1 bitmap path (matrices > 8 M columns).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spgemm.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r02c35_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c35_pytest.log
tail -3 gpurun_out/r02c35_pytest.log
