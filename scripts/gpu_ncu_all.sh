#!/bin/bash
# DRAM traffic + duration of the 30 fused SA-layer launches of one (warm) step, as CSV (small)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
   --clock-control none -k 'regex:sa_layer|thin_fwd|thin_bwd' -s 30 -c 30 --csv --log-file gpurun_out/ncu_all_layers.csv \
   python scripts/profile_sa.py all 2 > gpurun_out/ncu_all.log 2>&1
tail -1 gpurun_out/ncu_all.log; wc -l gpurun_out/ncu_all_layers.csv
