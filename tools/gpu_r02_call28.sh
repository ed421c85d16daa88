#!/bin/bash
# Round 2, GPU call 28 (1 GPU): numeric SpGEMM -- rows of 1024..8192 entries through the table kernel vs the dense kernel.
mkdir -p gpurun_out
for t in 8192 4096 2048 1024; do
  echo "B2S_SPGEMM_DENSE_MIN=$t" >> gpurun_out/r02c28_bench_spgemm.log
  B2S_SPGEMM_DENSE_MIN=$t timeout 300 python tools/bench_spgemm.py banded1m rmat16 rmat18 rmat20 >> gpurun_out/r02c28_bench_spgemm.log 2>&1
done
B2S_SPGEMM_DENSE_MIN=1024 timeout 300 python -m pytest tests/test_gpu_spgemm.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r02c28_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c28_pytest.log
grep "SPGEMM\|DENSE_MIN" gpurun_out/r02c28_bench_spgemm.log | cut -c1-200; tail -2 gpurun_out/r02c28_pytest.log
