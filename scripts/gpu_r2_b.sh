#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
bash scripts/gpu_ncu_fwd.sh
