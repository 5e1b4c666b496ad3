#!/bin/bash
# compute-sanitizer memcheck of one pytest selection: bash scripts/gpu_sanitize_one.sh <pytest -k expr> <file>
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --launch-timeout 600 --error-exitcode 0 --print-limit 20 \
  python -m pytest "$2" -q -x -k "$1" > gpurun_out/sanitize_one.log 2>&1
grep -n "Invalid\|Error\|at 0x\|by thread\|in b2r\|void b2r\|ERROR SUMMARY\|passed\|failed" gpurun_out/sanitize_one.log | head -60
