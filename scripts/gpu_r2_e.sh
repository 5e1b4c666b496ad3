#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest new"; timeout 900 python -m pytest tests/test_option_a_gpu.py tests/test_modules_gpu.py -q 2>&1 | tail -40 | tee gpurun_out/pytest_new.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
