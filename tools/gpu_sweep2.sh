mkdir -p gpurun_out
timeout 600 python tools/sweep_spmv.py l5 banded 2>&1 | grep -v "^#" | grep -E "waves  0|plan-free|copy"
