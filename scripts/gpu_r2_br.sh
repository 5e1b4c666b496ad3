#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_modules_gpu.py -m gpu -x -q -k "split_geometry or pipelined" 2>&1 | tail -3
for pa in 1 -1; do for fc in 4; do
timeout 600 python bench.py --workload br --no-cpu-baseline --steps 20 --warmup 5 --prepass-after $pa --fps-cluster $fc > gpurun_out/bench_brp.json 2> gpurun_out/bench_brp.err; tail -2 gpurun_out/bench_brp.err; python -c "
import json; d=json.load(open('gpurun_out/bench_brp.json')); print('br pa=$pa fc=$fc:', d['ms_per_step'], d['value'], d['e2e']['value'])"
done; done
timeout 600 python bench.py --workload br --no-cpu-baseline --steps 20 --warmup 5 --no-pipeline > gpurun_out/bench_brp.json 2> gpurun_out/bench_brp.err; python -c "
import json; d=json.load(open('gpurun_out/bench_brp.json')); print('br no-pipeline:', d['ms_per_step'], d['value'], d['e2e']['value'])"
