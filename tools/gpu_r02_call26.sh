#!/bin/bash
# Round 2, GPU call 26 (1 GPU): SpGEMM dense kernel: emission by one thread per level-0 word.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spgemm.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r02c26_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c26_pytest.log
timeout 300 python tools/bench_spgemm.py banded1m banded10m rmat16 rmat18 rmat20 > gpurun_out/r02c26_bench_spgemm.log 2>&1
B2S_BENCH_EXTRAS=spgemm_banded,spgemm_rmat timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/r02c26_bench_spgemm_extras.json 2> gpurun_out/r02c26_bench.err
tail -3 gpurun_out/r02c26_pytest.log; grep SPGEMM gpurun_out/r02c26_bench_spgemm.log; python -c "
import json; d=json.loads(open('gpurun_out/r02c26_bench_spgemm_extras.json').read()); print({k:v for k,v in d['extras'].items() if 'spgemm' in k})" 2>&1 | cut -c1-900
