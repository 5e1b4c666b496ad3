#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest ops"; timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_ref_pin_gpu.py tests/test_option_a_gpu.py -m gpu -x -q 2>&1 | tail -8
echo "== movers roofline"; timeout 600 python scripts/movers_roofline.py --json gpurun_out/movers_roofline.json 2>&1 | tee gpurun_out/movers_roofline.log
