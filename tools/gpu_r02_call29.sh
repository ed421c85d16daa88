#!/bin/bash
# Round 2, GPU call 29 (1 GPU): full GPU suite, bench + reference arm, SpMM table, launch list of the timed region.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02c29_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c29_pytest_gpu.log
timeout 600 python bench.py --steps 1000 --warmup 20 > gpurun_out/r02c29_bench_n1.json 2> gpurun_out/r02c29_bench_n1.err
timeout 200 python bench.py --impl reference --steps 50 --warmup 10 > gpurun_out/r02c29_bench_ref.json 2> gpurun_out/r02c29_bench_ref.err
timeout 300 python tools/bench_spmm.py 4000000 8 16 32 64 128 > gpurun_out/r02c29_bench_spmm.log 2>&1; cp gpurun_out/spmm_bench.json gpurun_out/r02c29_spmm_bench.json 2>/dev/null
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02c29_smoke.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 40 --csv --log-file gpurun_out/r02c29_launches_bench_n1.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/r02c29_launches_stdout.log 2>&1
tail -4 gpurun_out/r02c29_pytest_gpu.log; tail -2 gpurun_out/r02c29_smoke.log; head -c 1500 gpurun_out/r02c29_bench_n1.json; echo; head -c 600 gpurun_out/r02c29_bench_ref.json
