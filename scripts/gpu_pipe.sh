#!/bin/bash
# pipelined-step experiment: FPS width / caps around the default (pre-pass after SA2's forward)
mkdir -p gpurun_out
i=0
for cfg in "4|1|" "5|1|" "4|1|0:140,0:140,124:124,124:124,124:124" "5|1|0:140,0:140,116:116,116:116,116:116"; do
  i=$((i+1))
  IFS='|' read c after caps <<< "$cfg"
  python bench.py --no-cpu-baseline --steps 30 --fps-cluster $c --prepass-after $after ${caps:+--sm-caps $caps} > gpurun_out/bench_caps_$i.json 2> gpurun_out/bench_caps_$i.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_caps_$i.json"))
    print("fps $c after $after caps [$caps]", round(d["ms_per_step"], 4), round(d["value"], 1), round(d["e2e"]["value"], 1))
except Exception as e:
    print("fps $c after $after caps [$caps] failed", e)
PY
done
