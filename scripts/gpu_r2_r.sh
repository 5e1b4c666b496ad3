#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for e in 0 1; do
B2R_TORCH_ADAM=$e timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; tail -2 gpurun_out/bench_r.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r.json')); print('torch_adam=$e:', d['ms_per_step'], d['value'], d['e2e']['value'])"
done
