#!/bin/bash
# pipelined-step knobs after the pad-free position space: FPS cluster width x pre-pass start level
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
: > gpurun_out/sweep_r2.log
for pa in 1 0 -1; do
  for fc in 4 5 6 8; do
    timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 3 --fps-cluster $fc --prepass-after $pa > gpurun_out/_b.json 2> gpurun_out/_b.err
    python - "$fc" "$pa" >> gpurun_out/sweep_r2.log <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/_b.json"))
    print("fps_cluster %s prepass_after %s : %.3f ms/step  %.1f scenes/s  e2e %.1f" % (sys.argv[1], sys.argv[2], d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print("fps_cluster %s prepass_after %s : failed %s" % (sys.argv[1], sys.argv[2], e))
PY
  done
done
cat gpurun_out/sweep_r2.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 3 --trace gpurun_out/cupti_trace_r2.txt > gpurun_out/_b.json 2> gpurun_out/_b.err
head -70 gpurun_out/cupti_trace_r2.txt
