#!/bin/bash
# pipelined-step experiment: tests, then bench with different persistent-grid caps / FPS widths
mkdir -p gpurun_out
python -m pytest tests/test_modules_gpu.py -q -x > gpurun_out/pipe_tests.log 2>&1
tail -3 gpurun_out/pipe_tests.log
i=0
for cfg in "4|" "4|116:0,116:140,116:140,116:140,116:140" "5|" "5|108:0,108:140,108:140,108:140,108:140" "6|" "4|116:0,116:0,124:140,132:140,132:140"; do
  i=$((i+1))
  c=${cfg%%|*}; caps=${cfg#*|}
  python bench.py --no-cpu-baseline --steps 30 --fps-cluster $c ${caps:+--sm-caps $caps} > gpurun_out/bench_caps_$i.json 2> gpurun_out/bench_caps_$i.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_caps_$i.json"))
    print("cluster $c caps [$caps]", round(d["ms_per_step"], 4), round(d["value"], 1), round(d["e2e"]["value"], 1))
except Exception as e:
    print("cluster $c caps [$caps] failed", e)
PY
done
