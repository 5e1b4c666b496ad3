#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest fps"; timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k fps > gpurun_out/pytest_k.log 2>&1; tail -15 gpurun_out/pytest_k.log
echo "== sweep"; timeout 600 python scripts/fps_sweep.py 2>&1 | tee gpurun_out/fps_sweep_bucket.log
