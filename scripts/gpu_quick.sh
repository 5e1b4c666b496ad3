#!/bin/bash
# quick GPU pass: MLP + module parity tests, per-layer timing, bench without the CPU arms
mkdir -p gpurun_out
python -m pytest tests/test_mlp_gpu.py tests/test_modules_gpu.py -q -x > gpurun_out/quick_tests.log 2>&1; tail -3 gpurun_out/quick_tests.log
python scripts/time_sa.py > gpurun_out/time_sa_quick.log 2>&1; grep "sa[12] \|sum" gpurun_out/time_sa_quick.log
python bench.py --no-cpu-baseline --steps 30 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python -c "
import json
d = json.load(open('gpurun_out/bench_quick.json')); print('bench', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'])"
