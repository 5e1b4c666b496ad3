#!/bin/bash
# BR (BASELINE configs[2]) and GroupFree3D (configs[3]) steps at N GPUs, as the driver launches
# bench.py; short timeouts (a hang must not eat the budget)
N=${1:-2}
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for W in br gf3d; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --workload $W --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${W}_${N}gpu.json 2> gpurun_out/bench_${W}_${N}gpu.err
  echo "$W rc=$?"; grep -i "fail\|error\|retry" gpurun_out/bench_${W}_${N}gpu.err | head -5
  python -c "
import json
d = json.load(open('gpurun_out/bench_${W}_${N}gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['workload'][:50])"
done
