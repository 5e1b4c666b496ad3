#!/bin/bash
# ncu --set full + source of the SA1 top-layer and dense backward launches (second iteration)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sa_layer_bwd_kernel -s 2 -c 2 \
   -o gpurun_out/prof_two -f python scripts/profile_sa.py sa1 2 > gpurun_out/ncu_two.log 2>&1
tail -2 gpurun_out/ncu_two.log
