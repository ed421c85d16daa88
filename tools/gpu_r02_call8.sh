#!/bin/bash
# Round 2, GPU call 8 (1 GPU): SpGEMM with the hub-row class (tests + R-MAT scale 22 / 20 timings).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spgemm.py -x -q -p no:cacheprovider > gpurun_out/r02c8_pytest_spgemm.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c8_pytest_spgemm.log
B2S_BENCH_EXTRAS=spgemm_rmat timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/r02c8_bench_rmat22.json 2> gpurun_out/r02c8_bench_rmat22.err
timeout 300 python tools/bench_spgemm.py rmat18 rmat20 banded10m > gpurun_out/r02c8_bench_spgemm.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02c8_launches_spgemm_rmat18.csv python tools/bench_spgemm.py rmat18 > /dev/null 2>&1
tail -3 gpurun_out/r02c8_pytest_spgemm.log; tail -6 gpurun_out/r02c8_bench_spgemm.log; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02c8_bench_rmat22.json').read().split('\n') if l.startswith('{')][-1])
print(d.get('extras'))
PY
