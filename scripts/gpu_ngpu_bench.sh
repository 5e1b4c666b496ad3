#!/bin/bash
# N-GPU bench with the captured collective; short timeout (a hang must not eat the budget)
N=${1:-2}
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "rc=$?"; grep -i "fail\|error\|retry" gpurun_out/bench_${N}gpu.err | head -5
python -c "
import json
d = json.load(open('gpurun_out/bench_${N}gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['launch'][:60])"
