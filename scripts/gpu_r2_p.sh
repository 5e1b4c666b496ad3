#!/bin/bash
# N-GPU bench: captured collective vs eager collective (usage: gpu_r2_p.sh N)
N=${1:-2}
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu_$name.json 2> gpurun_out/bench_${N}gpu_$name.err
  echo "rc=$?"; grep -i "fail\|error\|retry" gpurun_out/bench_${N}gpu_$name.err | head -5
  python -c "
import json
d = json.load(open('gpurun_out/bench_${N}gpu_$name.json')); print('$name', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['launch'][:60])"
}
run captured B2R_X=0
run eager B2R_EAGER_COLLECTIVE=1
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_1gpu_ref.json 2> gpurun_out/bench_1gpu_ref.err
python -c "
import json
d = json.load(open('gpurun_out/bench_1gpu_ref.json')); print('1gpu', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
