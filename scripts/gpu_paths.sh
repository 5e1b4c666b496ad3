#!/bin/bash
# the other bench paths (eager, un-pipelined graph) and the GroupFree3D backbone configuration
mkdir -p gpurun_out
python scripts/gf3d_check.py > gpurun_out/gf3d_check.log 2>&1; tail -2 gpurun_out/gf3d_check.log
for f in "--no-pipeline" "--no-graph"; do
  python bench.py --no-cpu-baseline --steps 10 $f > gpurun_out/bench_alt.json 2> gpurun_out/bench_alt.err || tail -5 gpurun_out/bench_alt.err
  python -c "
import json
d = json.load(open('gpurun_out/bench_alt.json')); print('$f', d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['launch'][:40])"
done
