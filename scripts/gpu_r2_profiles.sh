#!/bin/bash
# round-2 evidence batch: tests, smoke, the three bench workloads, trace, per-layer timing, ncu launch
# list, ncu DRAM traffic of the SA layers, ncu sections of the FPS kernels, microbench vs the reference
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
echo "== bench default"; timeout 1200 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -3 $O/bench_1gpu.err; cut -c1-260 $O/bench_1gpu.json
echo "== bench br"; timeout 600 python bench.py --workload br --no-cpu-baseline > $O/bench_br.json 2> $O/bench_br.err; cut -c1-200 $O/bench_br.json
echo "== bench gf3d"; timeout 600 python bench.py --workload gf3d --no-cpu-baseline > $O/bench_gf3d.json 2> $O/bench_gf3d.err; cut -c1-200 $O/bench_gf3d.json
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-300 $O/bench_reference.json
echo "== trace"; timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --trace $O/cupti_trace_pipelined_step.txt > $O/_b.json 2> $O/_b.err; head -2 $O/cupti_trace_pipelined_step.txt
echo "== time_sa"; timeout 300 python scripts/time_sa.py > $O/time_sa.log 2>&1; tail -3 $O/time_sa.log
echo "== time_dense"; timeout 300 python scripts/time_dense.py > $O/time_dense.log 2>&1; tail -2 $O/time_dense.log
echo "== launch list"; bash scripts/gpu_launch_list.sh
echo "== ncu SA layers"; bash scripts/gpu_ncu_all.sh
echo "== ncu fps"; timeout 600 ncu --section SpeedOfLight --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats \
   --clock-control none -k 'regex:fps_' -s 5 -c 5 python scripts/fps_ncu_target.py > $O/ncu_fps_sections.txt 2>&1; grep -c "fps_" $O/ncu_fps_sections.txt
echo "== microbench"; timeout 600 python scripts/microbench.py --json $O/microbench.json > $O/microbench.log 2>&1; tail -5 $O/microbench.log
echo "== determinism"; timeout 300 python scripts/bwd_determinism.py > $O/bwd_determinism.log 2>&1; tail -4 $O/bwd_determinism.log
