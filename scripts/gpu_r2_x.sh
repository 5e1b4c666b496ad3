#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests/test_modules_gpu.py tests/test_mlp_gpu.py tests/test_dense_gpu.py -m gpu -x -q 2>&1 | tail -3
for e in 1 0; do
B2R_ZERO_ARENA=$e timeout 600 python bench.py --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; tail -2 gpurun_out/bench_x.err; python -c "
import json; d=json.load(open('gpurun_out/bench_x.json')); print('arena=$e:', d['ms_per_step'], d['value'], d['e2e']['value'])"
done
for w in br gf3d; do timeout 600 python bench.py --workload $w --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; tail -2 gpurun_out/bench_x.err; python -c "
import json; d=json.load(open('gpurun_out/bench_x.json')); print('$w:', d['ms_per_step'], d['value'], d['e2e']['value'])"; done
