#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest modules"; timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_option_a_gpu.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; tail -2 gpurun_out/bench_w.err; python -c "
import json; d=json.load(open('gpurun_out/bench_w.json')); print('run $i:', d['ms_per_step'], d['value'], d['e2e']['value'])"
done
timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --trace gpurun_out/cupti_trace_w.txt > gpurun_out/_b.json 2> gpurun_out/_b.err; head -2 gpurun_out/cupti_trace_w.txt
