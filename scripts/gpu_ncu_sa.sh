#!/bin/bash
# one ncu --set full capture of the fused SA-layer kernels (2nd iteration = warm)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
W=${1:-sa1}; N=${2:-6}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sa_layer -s $N -c $N \
   -o gpurun_out/prof_$W -f python scripts/profile_sa.py $W 2 > gpurun_out/ncu_$W.log 2>&1
tail -2 gpurun_out/ncu_$W.log
