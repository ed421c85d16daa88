#!/bin/bash
# Round 2, GPU call 4 (2 GPUs): dedicated pusher CTAs + .cg remote gathers (fused halo), column-block experiment, e2e
# pipeline stages.   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_r02_call4.sh'
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dist.py -x -q -p no:cacheprovider > gpurun_out/r02c4_pytest_dist.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c4_pytest_dist.log
run_n2() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 200)) bench.py --gpus 2 --steps 500 --warmup 20 --no-extras > gpurun_out/r02c4_bench_n2_$tag.json 2> gpurun_out/r02c4_bench_n2_$tag.err; }
run_n2 default B2S_X=0
run_n2 default2 B2S_X=0
run_n2 nowait B2S_FUSE_DEBUG=no-wait B2S_BENCH_NOVERIFY=1
(CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/bench_blocks.py 2 4 8 > gpurun_out/r02c4_bench_blocks.log 2>&1) &
for st in 16 8 4 2; do B2S_PIPE_CHUNKS=$st CUDA_VISIBLE_DEVICES=1 timeout 100 python tools/e2e_trace.py 2>/dev/null | tail -1 | sed "s/^/stages $st: /" >> gpurun_out/r02c4_e2e_stages.log; done
CUDA_VISIBLE_DEVICES=1 timeout 300 python -m pytest tests/test_gpu_spmv.py -q -p no:cacheprovider -x > gpurun_out/r02c4_pytest_spmv.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c4_pytest_spmv.log
wait
tail -3 gpurun_out/r02c4_pytest_dist.log; tail -3 gpurun_out/r02c4_pytest_spmv.log; grep -h '"value"' gpurun_out/r02c4_bench_n2_*.json | cut -c1-170; cat gpurun_out/r02c4_e2e_stages.log; cat gpurun_out/r02c4_bench_blocks.log
