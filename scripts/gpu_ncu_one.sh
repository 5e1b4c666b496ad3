#!/bin/bash
# ncu --set full + source of ONE launch: gpu_ncu_one.sh <profile_sa arg> <kernel regex> <skip>
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 \
   -o gpurun_out/prof_one -f python scripts/profile_sa.py $1 2 > gpurun_out/ncu_one.log 2>&1
tail -1 gpurun_out/ncu_one.log
