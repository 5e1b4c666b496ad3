#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
(cd scripts/probe && export B2R_FPS_BUCKET_SMALL=1 && for a in "40000 2048 8" "40000 2048 4" "40000 2048 10" "2048 1024 0"; do ./fps_phase_probe $a; done) | tee gpurun_out/fps_phase_probe.log
echo "== pytest fps"; timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k fps > gpurun_out/pytest_k.log 2>&1; tail -3 gpurun_out/pytest_k.log
B2R_FPS_BUCKET_SMALL=1 timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k fps 2>&1 | tail -2
echo "== sweep"; timeout 600 python scripts/fps_sweep.py 2>&1 | grep "bucket\|N=40000.*fps.cu.*auto" | tee gpurun_out/fps_sweep_bucket.log
for fc in 4 5; do
echo "== bench fc=$fc"; timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --fps-cluster $fc > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; tail -3 gpurun_out/bench_l.err; python -c "
import json; d=json.load(open('gpurun_out/bench_l.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['fps'])"
done
