#!/bin/bash
# Round 2, GPU call 20 (8 GPUs): final-code validation at 2/4/8 ranks (fused exchange, column blocks on tile shape 13)
# and the scaling numbers on one box.   gpurun --gpus 8 --timeout 1200 -- tools/gpu_r02_call20.sh
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r02c20_ngpus.txt
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_c_abi_sharded.py -x -q -p no:cacheprovider > gpurun_out/r02c20_pytest_dist.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c20_pytest_dist.log
port=29540
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 500 --warmup 20 > gpurun_out/r02c20_bench_n$n.json 2> gpurun_out/r02c20_bench_n$n.err
  port=$((port+1))
done
B2S_BENCH_EXTRAS=r32,cg timeout 200 python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/r02c20_bench_n1.json 2> gpurun_out/r02c20_bench_n1.err
tail -3 gpurun_out/r02c20_pytest_dist.log; for n in 8 4 2 1; do grep -h '"value"' gpurun_out/r02c20_bench_n$n.json | cut -c1-200; tail -2 gpurun_out/r02c20_bench_n$n.err; done
