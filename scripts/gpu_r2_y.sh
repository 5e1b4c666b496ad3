#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests/test_modules_gpu.py tests/test_dense_gpu.py -m gpu -x -q 2>&1 | tail -3
for e in 0 1 0 1; do
B2R_PINGPONG=$e timeout 600 python bench.py --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err; tail -2 gpurun_out/bench_y.err; python -c "
import json; d=json.load(open('gpurun_out/bench_y.json')); print('pingpong=$e:', d['ms_per_step'], d['value'], d['e2e']['value'])"
done
timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --trace gpurun_out/cupti_trace_y.txt > gpurun_out/_b.json 2> gpurun_out/_b.err; head -2 gpurun_out/cupti_trace_y.txt
