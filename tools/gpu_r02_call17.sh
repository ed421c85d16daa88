#!/bin/bash
# Round 2, GPU call 17 (1 GPU): PCIe duplex rate vs copy granularity (the ceiling of the staged host-vector product).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spmv.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r02c17_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c17_pytest.log
timeout 200 python tools/pcie_duplex_chunks.py > gpurun_out/r02c17_pcie_duplex_chunks.log 2>&1
timeout 100 python tools/pcie_duplex.py > gpurun_out/r02c17_pcie_duplex.json 2>&1
tail -3 gpurun_out/r02c17_pytest.log; cat gpurun_out/r02c17_pcie_duplex_chunks.log gpurun_out/r02c17_pcie_duplex.json
