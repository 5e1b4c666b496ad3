#!/bin/bash
# ncu --set full + source of the dense tcgen05 kernels at the fp2.l0 shape (M=8192, 512->256)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_gemm_kernel -s 20 -c 5 \
   -o gpurun_out/prof_dense -f python scripts/time_dense.py 1 > gpurun_out/ncu_dense.log 2>&1
tail -3 gpurun_out/ncu_dense.log
python scripts/ncu_summary.py gpurun_out/prof_dense.ncu-rep > gpurun_out/dense_summary.txt 2>&1
cat gpurun_out/dense_summary.txt
for i in 0 1 4; do
  ncu -i gpurun_out/prof_dense.ncu-rep --page source --csv --print-source sass --launch-skip $i --launch-count 1 > gpurun_out/dense_sass_$i.csv 2>/dev/null
  ncu -i gpurun_out/prof_dense.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $i --launch-count 1 > gpurun_out/dense_src_$i.csv 2>/dev/null
  python scripts/ncu_top_stalls.py gpurun_out/dense_sass_$i.csv 12 > gpurun_out/dense_stalls_$i.txt 2>&1
  python scripts/ncu_top_lines.py gpurun_out/dense_src_$i.csv 25 > gpurun_out/dense_lines_$i.txt 2>&1
  echo "---- launch $i"; head -3 gpurun_out/dense_stalls_$i.txt; cat gpurun_out/dense_lines_$i.txt
done
