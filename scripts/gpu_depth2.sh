#!/bin/bash
# two-batch-deep geometry pipeline (PipelinedTrainStep2) vs the default
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for pa in 1 0; do for fc in 4 5; do python bench.py --no-cpu-baseline --steps 30 --warmup 5 --pipeline-depth 2 --prepass-after $pa --fps-cluster $fc > gpurun_out/_b.json 2> gpurun_out/_b.err; python -c "
import json; d=json.load(open('gpurun_out/_b.json')); print('depth2 pa=$pa fc=$fc', d['ms_per_step'], d['value'], d['e2e']['value'])"; done; done
