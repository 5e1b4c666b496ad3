#!/bin/bash
# ncu --set full + source of SA1's dense and pooled forward launches and its top backward launch
# (second iteration, pad-free position space): where do the pooled epilogues wait?
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sa_layer_fwd_kernel -s 2 -c 2 \
   -o gpurun_out/prof_fwd -f python scripts/profile_sa.py sa1 2 > gpurun_out/ncu_fwd.log 2>&1
tail -2 gpurun_out/ncu_fwd.log
for i in 0 1; do
  ncu -i gpurun_out/prof_fwd.ncu-rep --page source --csv --print-source sass --launch-skip $i --launch-count 1 > gpurun_out/fwd_sass_$i.csv 2>/dev/null
  ncu -i gpurun_out/prof_fwd.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $i --launch-count 1 > gpurun_out/fwd_src_$i.csv 2>/dev/null
  python scripts/ncu_top_stalls.py gpurun_out/fwd_sass_$i.csv 30 > gpurun_out/fwd_stalls_$i.txt 2>&1
  python scripts/ncu_top_lines.py gpurun_out/fwd_src_$i.csv 30 > gpurun_out/fwd_lines_$i.txt 2>&1
done
python scripts/ncu_summary.py gpurun_out/prof_fwd.ncu-rep > gpurun_out/fwd_summary.txt 2>&1
cat gpurun_out/fwd_summary.txt | head -40
cat gpurun_out/fwd_stalls_1.txt | head -45
