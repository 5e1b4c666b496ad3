#!/bin/bash
# compute-sanitizer over the kernels with hand-rolled synchronisation (mbarrier pipelines, TMEM
# hand-offs, DSMEM/cluster FPS): memcheck on the per-kernel tests of the fused SA layers, the dense
# layers and the ops; racecheck + synccheck on a small selection that reaches every kernel family
# (template instantiations: gather / dense / top, padded and pad-free position space, thin layer,
# FWD / DGRAD / WGRAD dense kernels, cluster FPS, grid ball query).  Logs -> gpurun_out/sanitize_*.
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
run() {   # run <tool> <tag> <pytest args...>
  tool=$1; tag=$2; shift 2
  timeout 1500 compute-sanitizer --tool "$tool" --launch-timeout 900 --error-exitcode 0 --print-limit 5 \
      python -m pytest "$@" -q -x -p no:cacheprovider > gpurun_out/sanitize_${tool}_${tag}.log 2>&1
  echo "== $tool $tag: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_${tag}.log | tail -1)  [$(grep -E 'passed|failed' gpurun_out/sanitize_${tool}_${tag}.log | tail -1)]"
}
SEL_MLP="dense_layer_store_and_stats or gather_layer_matches or pool_epilogue or dense_layer_backward or gather_layer_backward or thin_first_layer or compact_dense_layer or compact_plan"
run memcheck mlp tests/test_mlp_gpu.py -k "$SEL_MLP or fused_block_single_scene or compact_block"
run memcheck dense tests/test_dense_gpu.py
run memcheck ops tests/test_ops_gpu.py -k "fps_small or fps_odd or fps_hole or fps_forced or fps_bucket or fps_legacy or ball_query_grid or ball_query_ragged or three_nn or group or gather or interp or scatter"
run memcheck adam tests/test_modules_gpu.py -k "flat_adam"
run synccheck mlp tests/test_mlp_gpu.py -k "$SEL_MLP"
run synccheck dense tests/test_dense_gpu.py -k "dense_forward or dense_backward"
run synccheck ops tests/test_ops_gpu.py -k "fps_small_levels or fps_forced_cluster or fps_hole or fps_bucket or fps_cluster_hint or scatter"
run racecheck mlp tests/test_mlp_gpu.py -k "dense_layer_store_and_stats or pool_epilogue or compact_dense_layer or thin_first_layer_forward"
run racecheck mlpbwd tests/test_mlp_gpu.py -k "dense_layer_backward or gather_layer_backward"
run racecheck dense tests/test_dense_gpu.py -k "dense_forward"
run racecheck ops tests/test_ops_gpu.py -k "fps_small_levels or fps_forced_cluster or fps_bucket or ball_query_ragged or scatter_plan_cache"
grep -h -A12 -E "Race reported|Invalid|Barrier error|Error:" gpurun_out/sanitize_*.log | head -80
