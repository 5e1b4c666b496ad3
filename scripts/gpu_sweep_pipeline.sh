#!/bin/bash
# pipelined-step knobs: pre-pass start level x FPS cluster width (usage: gpu_sweep_pipeline.sh [votenet|br|gf3d])
W=${1:-votenet}
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
: > gpurun_out/sweep_$W.log
for pa in 1 0 -1; do for fc in 4 5 6 8 10; do
timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 30 --warmup 5 --prepass-after $pa --fps-cluster $fc > gpurun_out/_b.json 2> gpurun_out/_b.err
python -c "
import json; d=json.load(open('gpurun_out/_b.json')); print('$W prepass_after $pa fps_cluster $fc: %.3f ms  %.1f scenes/s  e2e %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value']))" | tee -a gpurun_out/sweep_$W.log
done; done
