#!/bin/bash
# pipelined-step experiment: when the next batch's pre-pass starts, and the matching grid caps
mkdir -p gpurun_out
python -m pytest tests/test_modules_gpu.py -q -x -k "pipelined or captured or geometry" > gpurun_out/pipe_tests.log 2>&1
tail -2 gpurun_out/pipe_tests.log
i=0
for cfg in "-1|" "1|0:140,0:140,116:116,116:116,116:116" "1|0:0,0:140,116:116,116:116,116:116" "0|0:0,116:140,116:116,116:116,116:116" "1|0:140,0:116,116:116,116:116,116:116" "2|0:140,0:116,0:116,116:116,116:116"; do
  i=$((i+1))
  c=${cfg%%|*}; caps=${cfg#*|}
  python bench.py --no-cpu-baseline --steps 30 --prepass-after $c ${caps:+--sm-caps $caps} > gpurun_out/bench_caps_$i.json 2> gpurun_out/bench_caps_$i.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_caps_$i.json"))
    print("after $c caps [$caps]", round(d["ms_per_step"], 4), round(d["value"], 1), round(d["e2e"]["value"], 1))
except Exception as e:
    print("after $c caps [$caps] failed", e)
PY
done
