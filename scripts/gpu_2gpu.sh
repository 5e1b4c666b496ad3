#!/bin/bash
# 2-GPU bench (torchrun, as the driver launches it) + the prefetcher/prepack tests
mkdir -p gpurun_out
python -m pytest tests/test_modules_gpu.py -q -x -k "prefetcher or prepacked" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -3 gpurun_out/bench_2gpu.err
python -c "
import json
d = json.load(open('gpurun_out/bench_2gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
