#!/bin/bash
# ncu --set full --import-source on of the dominant kernel class (round 2): the six SA1 launches of
# one fused block iteration (thin fwd, dense fwd, pooled fwd, top bwd, dense bwd, thin bwd)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:sa_layer_bwd|thin_fwd|thin_bwd|sa_layer_fwd' -s 6 -c 6 \
   -o gpurun_out/prof_sa1_r02 -f python scripts/profile_sa.py sa1 2 > gpurun_out/ncu_sa1_r02.log 2>&1
tail -1 gpurun_out/ncu_sa1_r02.log; ls -la gpurun_out/prof_sa1_r02.ncu-rep
