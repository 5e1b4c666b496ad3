#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest fps"; timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k fps 2>&1 | tail -3
echo "== sweep"; timeout 600 python scripts/fps_sweep.py 2>&1 | grep "N=2048\|N=1024\|N=512" | tee gpurun_out/fps_sweep_small.log
for depth in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --pipeline-depth $depth > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; tail -2 gpurun_out/bench_o.err; python -c "
import json; d=json.load(open('gpurun_out/bench_o.json')); print('depth $depth:', d['ms_per_step'], d['value'], d['e2e']['value'])"
done
