#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest modules"; timeout 900 python -m pytest tests/test_modules_gpu.py -m gpu -x -q > gpurun_out/pytest_j.log 2>&1; tail -3 gpurun_out/pytest_j.log
for pa in 1 0; do for fc in 4 5; do
echo "== bench pa=$pa fc=$fc"; timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --prepass-after $pa --fps-cluster $fc > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; tail -3 gpurun_out/bench_j.err; python -c "
import json; d=json.load(open('gpurun_out/bench_j.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'])"
done; done
timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --trace gpurun_out/cupti_trace_j.txt > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err
head -3 gpurun_out/cupti_trace_j.txt
