#!/bin/bash
# round 2, call A: the pad-free position space (csrc/compact.cu) -- parity, then step time with / without
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
echo "== pytest test_mlp_gpu (compact)"; timeout 900 python -m pytest tests/test_mlp_gpu.py -q -k "compact" 2>&1 | tail -40 | tee gpurun_out/pytest_compact.log
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== bench compact"; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_compact.json 2> gpurun_out/bench_compact.err; tail -3 gpurun_out/bench_compact.err; cut -c1-300 gpurun_out/bench_compact.json
echo "== bench padded"; B2R_COMPACT=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_padded.json 2> gpurun_out/bench_padded.err; tail -3 gpurun_out/bench_padded.err; cut -c1-300 gpurun_out/bench_padded.json
echo "== per-layer timing (compact)"; timeout 300 python scripts/time_sa.py > gpurun_out/time_sa_compact.log 2>&1; tail -30 gpurun_out/time_sa_compact.log
