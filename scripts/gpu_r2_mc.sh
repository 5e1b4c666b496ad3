#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "scatter or group or interpolate" 2>&1 | tail -2
for r in 4 8; do echo "== rows $r"; B2R_MC_ROWS=$r timeout 600 python scripts/movers_roofline.py 2>&1 | grep "group_bwd\|interp_bwd" | grep -v "N=40000\|N=100000\|N=20000\|B=1 "; done
