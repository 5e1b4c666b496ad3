#!/bin/bash
# round-2 re-entry baseline: full GPU tests, default bench + timeline trace
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
echo "== bench + trace"; timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --trace gpurun_out/cupti_trace_i.txt > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; tail -3 gpurun_out/bench_i.err; cut -c1-200 gpurun_out/bench_i.json
head -5 gpurun_out/cupti_trace_i.txt
