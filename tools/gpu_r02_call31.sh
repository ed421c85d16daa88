#!/bin/bash
# Round 2, GPU call 31 (1 GPU): column-split SpMV (spmv_domain_part, automatic for scattered matrices with x > L2).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_examples.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r02c31_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c31_pytest.log
B2S_BENCH_EXTRAS=r32,gmg timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/r02c31_bench.json 2> gpurun_out/r02c31_bench.err
tail -3 gpurun_out/r02c31_pytest.log; python -c "
import json; d=json.loads(open('gpurun_out/r02c31_bench.json').read()); 
for k,v in d['extras'].items(): print(k, json.dumps(v)[:900])"
