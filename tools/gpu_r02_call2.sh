#!/bin/bash
# Round 2, GPU call 2 (2 GPUs): the fused exchange (halo + column blocks), CG, C-ABI NCCL path, new 1-GPU tests,
# bench at N=2.   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_r02_call2.sh'
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02c2_gpus.txt 2>&1
timeout 500 python -m pytest tests/test_gpu_dist.py tests/test_gpu_c_abi_sharded.py -x -q -p no:cacheprovider > gpurun_out/r02c2_pytest_dist.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c2_pytest_dist.log
timeout 400 python -m pytest tests/test_gpu_zzassembly.py tests/test_gpu_spgemm.py tests/test_gpu_spmv.py -q -p no:cacheprovider -k "library_assembly or chunked or plan_kernel_choice or inplace or fused_entry or transpose or tocsr or device" > gpurun_out/r02c2_pytest_new.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c2_pytest_new.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 500 --warmup 20 > gpurun_out/r02c2_bench_n2.json 2> gpurun_out/r02c2_bench_n2.err
B2S_PEER_FUSED=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 500 --warmup 20 --no-extras > gpurun_out/r02c2_bench_n2_nccl.json 2> gpurun_out/r02c2_bench_n2_nccl.err
timeout 60 python tools/pcie_duplex.py > gpurun_out/r02c2_pcie.json 2>&1
SWEEP_CFGS=0,6,11 timeout 120 python tools/sweep_spmv.py banded > gpurun_out/r02c2_sweep_banded.log 2>&1
tail -4 gpurun_out/r02c2_pytest_dist.log; tail -4 gpurun_out/r02c2_pytest_new.log; head -c 1200 gpurun_out/r02c2_bench_n2.json; tail -5 gpurun_out/r02c2_bench_n2.err; cat gpurun_out/r02c2_pcie.json
