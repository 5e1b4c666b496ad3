#!/bin/bash
# racecheck of the fused backward kernels only (see scripts/gpu_sanitize.sh for the full set)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --launch-timeout 900 --error-exitcode 0 --print-limit 5 \
   python -m pytest tests/test_mlp_gpu.py -k "dense_layer_backward or gather_layer_backward" -q -x -p no:cacheprovider > gpurun_out/sanitize_racecheck_mlpbwd.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_racecheck_mlpbwd.log | tail -2
grep -h "Race reported" gpurun_out/sanitize_racecheck_mlpbwd.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | head
python -m pytest tests/test_mlp_gpu.py -m gpu -x -q 2>&1 | tail -1
python scripts/time_sa.py 2>&1 | grep "sa_layer_bwd:\|sum of"
