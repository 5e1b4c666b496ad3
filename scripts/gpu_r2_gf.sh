#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for pa in 1 0; do for fc in 4 6 8 10; do
timeout 600 python bench.py --workload gf3d --no-cpu-baseline --steps 30 --warmup 5 --fps-cluster $fc --prepass-after $pa > gpurun_out/bench_gfp.json 2> gpurun_out/bench_gfp.err; tail -2 gpurun_out/bench_gfp.err; python -c "
import json; d=json.load(open('gpurun_out/bench_gfp.json')); print('gf3d pa=$pa fc=$fc:', d['ms_per_step'], d['value'], d['e2e']['value'])"
done; done
