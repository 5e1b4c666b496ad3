#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest pipelined"; timeout 900 python -m pytest tests/test_modules_gpu.py -m gpu -x -q -k "pipelined" 2>&1 | tail -3
for depth in 2 1; do for pa in 1 0; do for fc in 4 5; do
timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --pipeline-depth $depth --prepass-after $pa --fps-cluster $fc > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; tail -2 gpurun_out/bench_m.err; python -c "
import json; d=json.load(open('gpurun_out/bench_m.json')); print('depth $depth pa $pa fc $fc:', d['ms_per_step'], d['value'], d['e2e']['value'])"
done; done; done
timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --pipeline-depth 2 --prepass-after 1 --fps-cluster 4 --trace gpurun_out/cupti_trace_m.txt > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err
head -3 gpurun_out/cupti_trace_m.txt
