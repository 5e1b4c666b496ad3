#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for fc in 4 5 6 8; do
timeout 600 python bench.py --workload br --no-cpu-baseline --steps 20 --warmup 5 --fps-cluster $fc > gpurun_out/bench_brp.json 2> gpurun_out/bench_brp.err; tail -2 gpurun_out/bench_brp.err; python -c "
import json; d=json.load(open('gpurun_out/bench_brp.json')); print('br fc=$fc:', d['ms_per_step'], d['value'], d['e2e']['value'])"
done
