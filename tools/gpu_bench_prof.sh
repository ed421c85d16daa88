mkdir -p gpurun_out
timeout 600 python bench.py --steps 1000 --warmup 20 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/launches_stdout.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_tma -s 5 -c 1 -f -o gpurun_out/prof_r01_spmv_l5 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/prof_stdout.log 2>&1
ls -la gpurun_out
