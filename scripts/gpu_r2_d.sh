#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "== pytest pipelined"; timeout 900 python -m pytest tests/test_modules_gpu.py -q -k "pipelined or captured" 2>&1 | tail -15 | tee gpurun_out/pytest_pipe.log
: > gpurun_out/sweep_d2.log
for depth in 2 1; do
 for pa in 1 0 -1; do
  for fc in 4 5 8; do
    timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 3 --pipeline-depth $depth --fps-cluster $fc --prepass-after $pa > gpurun_out/_b.json 2> gpurun_out/_b.err
    python - "$depth" "$fc" "$pa" >> gpurun_out/sweep_d2.log <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/_b.json"))
    print("depth %s fps_cluster %s prepass_after %s : %.3f ms/step  %.1f scenes/s  e2e %.1f" % (sys.argv[1], sys.argv[2], sys.argv[3], d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print("depth %s fps_cluster %s prepass_after %s : failed %s" % (sys.argv[1], sys.argv[2], sys.argv[3], e)); print(open("gpurun_out/_b.err").read()[-600:])
PY
  done
 done
done
cat gpurun_out/sweep_d2.log
