#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_gpu.py -x -q 2>&1 | tail -40 | tee gpurun_out/mlp_test.log
