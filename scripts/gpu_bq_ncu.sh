#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k ball_query 2>&1 | tail -2
timeout 300 python scripts/microbench.py 2>&1 | grep -E "ball_query" | tee gpurun_out/microbench_bq.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:grid_ -c 12 --csv --log-file gpurun_out/bq_grid.csv python scripts/microbench.py > /dev/null 2>&1
grep '^"' gpurun_out/bq_grid.csv | cut -d, -f5,15 | head -14
