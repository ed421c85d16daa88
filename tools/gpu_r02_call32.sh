#!/bin/bash
# Round 2, GPU call 32 (1 GPU): full GPU suite, bench + reference arm, SpMM table, launch list of the timed region.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02c32_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c32_pytest_gpu.log
timeout 600 python bench.py --steps 1000 --warmup 20 > gpurun_out/r02c32_bench_n1.json 2> gpurun_out/r02c32_bench_n1.err
timeout 200 python bench.py --impl reference --steps 50 --warmup 10 > gpurun_out/r02c32_bench_ref.json 2> gpurun_out/r02c32_bench_ref.err
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02c32_smoke.log 2>&1
tail -4 gpurun_out/r02c32_pytest_gpu.log; tail -2 gpurun_out/r02c32_smoke.log; head -c 1500 gpurun_out/r02c32_bench_n1.json; echo; head -c 600 gpurun_out/r02c32_bench_ref.json
