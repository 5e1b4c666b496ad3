#!/bin/bash
# thin first-layer kernels: parity tests, per-layer timing, bench
mkdir -p gpurun_out
python -m pytest tests/test_mlp_gpu.py -q -x > gpurun_out/thin_tests.log 2>&1; tail -3 gpurun_out/thin_tests.log
python scripts/time_sa.py > gpurun_out/time_sa_thin.log 2>&1; grep "sa1\|sum" gpurun_out/time_sa_thin.log
python bench.py --no-cpu-baseline --steps 30 > gpurun_out/bench_thin.json 2> gpurun_out/bench_thin.err
python -c "
import json
d = json.load(open('gpurun_out/bench_thin.json')); print('thin', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'])"
