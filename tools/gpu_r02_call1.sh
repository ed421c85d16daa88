#!/bin/bash
# Round 2, GPU call 1 (1 GPU): parity of the new kernel paths, the gather ceiling, sweeps, bench with extras, ncu.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02_call1.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_gpu.txt 2>&1
timeout 120 tools/bin/gather_bench > gpurun_out/r02_gather_bench.txt 2>&1; echo "gather exit $?" >> gpurun_out/r02_gather_bench.txt
timeout 700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log
rm -f gpurun_out/sweep_spmv.txt
SWEEP_FLAVORS=0,2 SWEEP_CFGS=0,3,4,5,6,10,11 timeout 200 python tools/sweep_spmv.py l5 l5f32 > gpurun_out/r02_sweep_l5.log 2>&1
SWEEP_CFGS=0,3,4,5 timeout 120 python tools/sweep_spmv.py banded b32f32 >> gpurun_out/r02_sweep_l5.log 2>&1
SWEEP_CFGS=4,7,8,9,3 timeout 200 python tools/sweep_spmv.py r32 r32f64 > gpurun_out/r02_sweep_r32.log 2>&1
cp gpurun_out/sweep_spmv.txt gpurun_out/r02_sweep_spmv.txt
timeout 600 python bench.py --steps 1000 --warmup 20 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 200 python bench.py --impl reference --steps 50 --warmup 10 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_tma -s 3 -c 1 -f -o gpurun_out/r02_prof_spmv_l5 python tools/prof_spmv.py l5 -1 0 5 > gpurun_out/r02_prof_l5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_tma -s 3 -c 1 -f -o gpurun_out/r02_prof_spmv_r32 python tools/prof_spmv.py r32 -1 0 5 > gpurun_out/r02_prof_r32.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_bench_n1.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_launches_stdout.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log; tail -5 gpurun_out/r02_sweep_l5.log; head -c 1500 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
