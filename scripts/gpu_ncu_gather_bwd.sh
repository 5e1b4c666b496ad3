#!/bin/bash
# ncu --set full of the SA3 first-layer (Cin = 259 gather layer, 32-position tiles) backward and forward launches
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sa_layer_bwd_kernel -s 21 -c 1 \
   -o gpurun_out/prof_sa3_l0_bwd -f python scripts/profile_sa.py all 2 > gpurun_out/ncu_sa3b.log 2>&1
tail -1 gpurun_out/ncu_sa3b.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sa_layer_fwd_kernel -s 18 -c 1 \
   -o gpurun_out/prof_sa3_l0_fwd -f python scripts/profile_sa.py all 2 > gpurun_out/ncu_sa3f.log 2>&1
tail -1 gpurun_out/ncu_sa3f.log
