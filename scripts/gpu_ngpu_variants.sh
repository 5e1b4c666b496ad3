#!/bin/bash
# N-GPU bench under a few collective settings (usage: gpu_ngpu_variants.sh N)
N=${1:-4}
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu_$name.json 2> gpurun_out/bench_${N}gpu_$name.err
  python -c "
import json
d = json.load(open('gpurun_out/bench_${N}gpu_$name.json')); print('$name', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])" 2>&1 | tail -1
}
run default B2R_X=0
run overlap B2R_OVERLAP=1
run ll128 NCCL_PROTO=LL128
run tree NCCL_ALGO=Tree
run nvls NCCL_ALGO=NVLS
