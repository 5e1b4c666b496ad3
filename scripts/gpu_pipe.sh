#!/bin/bash
# pipelined-step experiment: tests, then bench with different persistent-grid caps
mkdir -p gpurun_out
python -m pytest tests/test_modules_gpu.py -q -x -k "pipelined or captured or geometry_stream" > gpurun_out/pipe_tests.log 2>&1
tail -3 gpurun_out/pipe_tests.log
i=0
for caps in "" "116:0,116:140,116:140,116:140,116:140" "116:140,116:140,116:140,116:140,116:140" "116:0,116:0,116:0,116:0,116:0" "116:0,116:0,140:140,140:140,140:140" "112:0,112:0,112:140,112:140,112:140" "116:116,116:140,116:140,116:140,116:140"; do
  i=$((i+1))
  python bench.py --no-cpu-baseline --steps 30 ${caps:+--sm-caps $caps} > gpurun_out/bench_caps_$i.json 2> gpurun_out/bench_caps_$i.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_caps_$i.json"))
    print("caps [$caps]", round(d["ms_per_step"], 4), round(d["value"], 1), round(d["e2e"]["value"], 1))
except Exception as e:
    print("caps [$caps] failed", e)
PY
done
