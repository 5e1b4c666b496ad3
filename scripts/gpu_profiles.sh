#!/bin/bash
# refresh profiles/: DRAM traffic of every fused SA-layer launch, ncu launch list of one eager
# step, ncu --set full of the thin first-layer kernels + the SA1 dense / top backward launches
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
bash scripts/gpu_ncu_all.sh
bash scripts/gpu_launch_list.sh
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:sa_layer_bwd|thin_fwd|thin_bwd|sa_layer_fwd' -s 6 -c 6 \
   -o gpurun_out/prof_sa1_r01b -f python scripts/profile_sa.py sa1 2 > gpurun_out/ncu_sa1b.log 2>&1
tail -1 gpurun_out/ncu_sa1b.log
