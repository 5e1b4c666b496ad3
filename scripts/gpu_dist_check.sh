#!/bin/bash
N=${1:-2}
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_flat_adam_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3
for ov in 0 1; do
B2R_NO_OVERLAP=$ov timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu_ov$ov.json 2> gpurun_out/bench_${N}gpu_ov$ov.err
echo "no_overlap=$ov rc=$?"; grep -i "fail\|error\|retry" gpurun_out/bench_${N}gpu_ov$ov.err | head -3
python -c "
import json
d = json.load(open('gpurun_out/bench_${N}gpu_ov$ov.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_1gpu_ref.json 2> gpurun_out/bench_1gpu_ref.err
python -c "
import json
d = json.load(open('gpurun_out/bench_1gpu_ref.json')); print('1gpu', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
