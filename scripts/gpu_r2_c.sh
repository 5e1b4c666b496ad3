#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest dense"; timeout 900 python -m pytest tests/test_dense_gpu.py -q 2>&1 | tail -40 | tee gpurun_out/pytest_dense.log
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== bench dense"; timeout 600 python bench.py --no-cpu-baseline --trace gpurun_out/cupti_trace_dense.txt > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err; tail -3 gpurun_out/bench_dense.err; cut -c1-300 gpurun_out/bench_dense.json
echo "== bench torch heads"; B2R_DENSE=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_nodense.json 2> gpurun_out/bench_nodense.err; tail -3 gpurun_out/bench_nodense.err; cut -c1-300 gpurun_out/bench_nodense.json
head -60 gpurun_out/cupti_trace_dense.txt
