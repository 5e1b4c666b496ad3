#!/bin/bash
# A/B of two builds of libb2r on one box: bash scripts/gpu_ab_lib.sh <other .so relative to the repo>
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
OTHER=$GRAFT_REPO_ROOT/$1
for rep in 1 2; do for lib in "" "$OTHER"; do
B2R_LIB=$lib python scripts/time_sa.py 2>&1 | grep "sa_layer_bwd:\|sum of" | tr '\n' ' '; echo " [lib=${lib:-default}]"
B2R_LIB=$lib python bench.py --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/_b.json 2> gpurun_out/_b.err; python -c "
import json; d=json.load(open('gpurun_out/_b.json')); print('   bench', d['ms_per_step'], d['value'], d['e2e']['value'])"
done; done
