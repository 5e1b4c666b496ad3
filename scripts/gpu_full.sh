#!/bin/bash
# full GPU parity suite, smoke, the default bench line, per-layer timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python -c "
import json
d = json.load(open('gpurun_out/bench.json')); print('bench', d['ms_per_step'], d['value'], d['e2e'], d['roofline']['frac'], d.get('reference_gpu', {}).get('value'), d.get('cpu_baseline', {}).get('value'))"
timeout 300 python scripts/time_sa.py > gpurun_out/time_sa.log 2>&1; tail -1 gpurun_out/time_sa.log
