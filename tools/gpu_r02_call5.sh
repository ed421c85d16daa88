#!/bin/bash
# Round 2, GPU call 5 (2 GPUs): gather mode for all-gather matrices, scattered-short configs, graded e2e pipeline,
# full N=2 bench with extras.   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_r02_call5.sh'
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dist.py -x -q -p no:cacheprovider > gpurun_out/r02c5_pytest_dist.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c5_pytest_dist.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 500 --warmup 20 > gpurun_out/r02c5_bench_n2.json 2> gpurun_out/r02c5_bench_n2.err
(CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_cg.py tests/test_gpu_zspmm.py -q -p no:cacheprovider -x > gpurun_out/r02c5_pytest_1gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c5_pytest_1gpu.log) &
for st in 0 6 4; do B2S_PIPE_CHUNKS=$st CUDA_VISIBLE_DEVICES=1 timeout 100 python tools/e2e_trace.py 2>/dev/null | tail -1 | sed "s/^/stages $st (0 = graded default): /" >> gpurun_out/r02c5_e2e_stages.log; done
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --steps 1000 --warmup 20 --no-cpu > gpurun_out/r02c5_bench_n1.json 2> gpurun_out/r02c5_bench_n1.err
wait
tail -3 gpurun_out/r02c5_pytest_dist.log; tail -3 gpurun_out/r02c5_pytest_1gpu.log; cat gpurun_out/r02c5_e2e_stages.log; grep -h '"value"' gpurun_out/r02c5_bench_n2.json | cut -c1-200; tail -3 gpurun_out/r02c5_bench_n2.err
