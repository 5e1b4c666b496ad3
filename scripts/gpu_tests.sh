#!/bin/bash
# parity tests on the GPU box; each file under its own timeout so a hung kernel cannot eat the lease
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for f in ${@:-tests/test_mlp_gpu.py tests/test_modules_gpu.py}; do
  echo "== $f"; timeout 900 python -m pytest $f -x -q 2>&1 | tail -30
done | tee gpurun_out/gpu_tests.log
