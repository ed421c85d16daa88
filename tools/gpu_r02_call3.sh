#!/bin/bash
# Round 2, GPU call 3 (2 GPUs): re-validate after the short-row threshold / push-placement changes, fused-halo timing
# experiments, SpMM window kernel, e2e pipeline trace.  gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_r02_call3.sh'
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dist.py -x -q -p no:cacheprovider > gpurun_out/r02c3_pytest_dist.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c3_pytest_dist.log
run_n2() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 200)) bench.py --gpus 2 --steps 500 --warmup 20 --no-extras > gpurun_out/r02c3_bench_n2_$tag.json 2> gpurun_out/r02c3_bench_n2_$tag.err; }
run_n2 default B2S_X=0
run_n2 orderonly B2S_FUSE_DEBUG=order-only B2S_BENCH_NOVERIFY=1
run_n2 nowait B2S_FUSE_DEBUG=no-wait B2S_BENCH_NOVERIFY=1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 500 --warmup 20 > gpurun_out/r02c3_bench_n2_full.json 2> gpurun_out/r02c3_bench_n2_full.err
# GPU 0: the whole 1-GPU suite; GPU 1 meanwhile: SpMM kernels
(CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_dist.py > gpurun_out/r02c3_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c3_pytest_gpu.log) &
CUDA_VISIBLE_DEVICES=1 timeout 200 python tools/bench_spmm.py 4000000 8 32 128 > gpurun_out/r02c3_bench_spmm.log 2>&1
cp gpurun_out/spmm_bench.json gpurun_out/r02c3_spmm_bench.json 2>/dev/null
wait
B2S_PIPE_TRACE=1 CUDA_VISIBLE_DEVICES=1 timeout 100 python tools/e2e_trace.py > gpurun_out/r02c3_e2e_trace.log 2>&1
tail -4 gpurun_out/r02c3_pytest_dist.log; tail -6 gpurun_out/r02c3_pytest_gpu.log; grep -h '"value"' gpurun_out/r02c3_bench_n2_*.json | cut -c1-200; tail -25 gpurun_out/r02c3_e2e_trace.log; grep SPMM gpurun_out/r02c3_bench_spmm.log | cut -c1-300
