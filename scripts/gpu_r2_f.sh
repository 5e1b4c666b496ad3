#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== bench default"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -4 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
echo "== bench br"; timeout 900 python bench.py --workload br --no-cpu-baseline > gpurun_out/bench_br.json 2> gpurun_out/bench_br.err; tail -4 gpurun_out/bench_br.err; cut -c1-300 gpurun_out/bench_br.json
echo "== bench gf3d"; timeout 900 python bench.py --workload gf3d --no-cpu-baseline > gpurun_out/bench_gf3d.json 2> gpurun_out/bench_gf3d.err; tail -4 gpurun_out/bench_gf3d.err; cut -c1-300 gpurun_out/bench_gf3d.json
echo "== time_sa"; timeout 300 python scripts/time_sa.py > gpurun_out/time_sa.log 2>&1; tail -36 gpurun_out/time_sa.log
