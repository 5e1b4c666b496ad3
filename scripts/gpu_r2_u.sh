#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests/test_modules_gpu.py tests/test_mlp_gpu.py tests/test_dense_gpu.py -m gpu -x -q 2>&1 | tail -4
for e in 0 1; do
B2R_NO_PINGPONG=$e timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err; tail -2 gpurun_out/bench_u.err; python -c "
import json; d=json.load(open('gpurun_out/bench_u.json')); print('no_pingpong=$e:', d['ms_per_step'], d['value'], d['e2e']['value'])"
done
