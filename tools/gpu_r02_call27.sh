#!/bin/bash
# Round 2, GPU call 27 (1 GPU): where SpGEMM's time goes on R-MAT after the expansion-loop changes: launch list at
# scale 20, ncu sections of the dense numeric / symbolic launches at scale 18.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02c27_launches_spgemm_rmat20.csv python tools/bench_spgemm.py rmat20 > gpurun_out/r02c27_launches_stdout.log 2>&1
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section LaunchStats --section SchedulerStats --clock-control none -k regex:spgemm_dense_kernel -c 2 -f -o gpurun_out/r02c27_prof_spgemm_dense python tools/bench_spgemm.py rmat18 > gpurun_out/r02c27_prof_spgemm.log 2>&1
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section LaunchStats --section SchedulerStats --clock-control none -k regex:spgemm_cta_kernel -s 2 -c 2 -f -o gpurun_out/r02c27_prof_spgemm_cta python tools/bench_spgemm.py rmat18 > gpurun_out/r02c27_prof_spgemm_cta.log 2>&1
tail -3 gpurun_out/r02c27_prof_spgemm.log; tail -3 gpurun_out/r02c27_launches_stdout.log
