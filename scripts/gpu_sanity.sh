#!/bin/bash
# first GPU contact: op parity tests + per-op timing vs the reference's own kernels
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests/test_ops_gpu.py -x -q 2>&1 | tail -40 | tee gpurun_out/ops_test.log
timeout 600 python scripts/microbench.py --json gpurun_out/microbench.json 2>&1 | tee gpurun_out/microbench.log
