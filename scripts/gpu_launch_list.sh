#!/bin/bash
# ncu launch list of ~one eager step (shares per kernel; absolute times are cold/serialised)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 1000 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches.csv
