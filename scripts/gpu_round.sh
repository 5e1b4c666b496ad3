#!/bin/bash
# full GPU pass: parity tests, smoke, bench (both arms), per-layer timing, ncu launch list
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench (product arm)"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-300 gpurun_out/bench_reference.json
echo "== per-layer timing"; timeout 300 python scripts/time_sa.py > gpurun_out/time_sa.log 2>&1; tail -1 gpurun_out/time_sa.log
echo "== ncu launch list"; bash scripts/gpu_launch_list.sh
