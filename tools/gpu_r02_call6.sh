#!/bin/bash
# Round 2, GPU call 6 (8 GPUs): multi-GPU tests at 2/4/8 ranks, bench at N=8 and N=4 with extras.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_r02_call6.sh'
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r02c6_ngpus.txt
timeout 400 python -m pytest tests/test_gpu_dist.py -x -q -p no:cacheprovider -k "nccl and 8" > gpurun_out/r02c6_pytest_dist.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c6_pytest_dist.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 500 --warmup 20 > gpurun_out/r02c6_bench_n8.json 2> gpurun_out/r02c6_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 500 --warmup 20 > gpurun_out/r02c6_bench_n4.json 2> gpurun_out/r02c6_bench_n4.err
timeout 120 python bench.py --steps 500 --warmup 20 --no-cpu --no-extras > gpurun_out/r02c6_bench_n1.json 2> gpurun_out/r02c6_bench_n1.err
tail -3 gpurun_out/r02c6_pytest_dist.log; for n in 8 4 1; do grep -h '"value"' gpurun_out/r02c6_bench_n$n.json | cut -c1-230; tail -2 gpurun_out/r02c6_bench_n$n.err; done
