#!/bin/bash
# Round 2, GPU call 18 (1 GPU): stage patterns and copy alignment of the host-vector pipeline.
mkdir -p gpurun_out
timeout 300 python tools/e2e_patterns.py > gpurun_out/r02c18_e2e_patterns.log 2>&1
cat gpurun_out/r02c18_e2e_patterns.log
