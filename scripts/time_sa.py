"""CUDA-event timing of every fused SA-layer launch of one VoteNet step (B = 8, 40k-point room
scenes, the real FPS-sampled chain sa1 -> sa2 -> sa3 -> sa4 -> vote aggregation):
   python scripts/time_sa.py [iters]
-> per launch: microseconds, algorithmic TFLOP/s (SURVEY 8d flops of the padded computation) and
fraction of the tensor peak of its operand type, block-level compulsory GB/s.  The step runs
eagerly (events cannot be recorded inside a graph replay); L2 is flushed between iterations."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backtoreality_b200 import _ext, scenes  # noqa: E402
from backtoreality_b200.votenet import VoteNet  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda:0")
torch.manual_seed(0)
B = 8
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
bf16 = float(peaks.get("bf16_tflops_sustained", 1400.0))
hbm = float(peaks.get("hbm_gbs", 6650.0))
net = VoteNet(22, 1, 22, np.ones((22, 3), np.float32), input_feature_dim=1, num_proposal=256).to(dev).train()
pcs = [torch.from_numpy(scenes.batch(1000 + 8 * i, B, 40000, C=1, kind="room", dup=0.2)).to(dev) for i in range(2)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
_ext.TIME_OPS.update(["sa_layer_fwd", "sa_layer_bwd"])
for it in range(iters + 2):
    if it == 2:
        torch.cuda.synchronize()
        _ext.TIMED.clear()
    flush.zero_()
    for p in net.parameters():
        p.grad = None
    ep = net({"point_clouds": pcs[it % 2]})
    ((ep["proposal_scores_raw"] ** 2).mean() + ((ep["vote_xyz"] - ep["seed_xyz"]) ** 2).mean()).backward()
torch.cuda.synchronize()
blocks = ["sa1", "sa2", "sa3", "sa4", "vote"]
total = {"sa_layer_fwd": 0.0, "sa_layer_bwd": 0.0}
flops = {"sa_layer_fwd": 0.0, "sa_layer_bwd": 0.0}
for op in ("sa_layer_fwd", "sa_layer_bwd"):
    ev = _ext.TIMED.get(op, [])
    per = len(ev) // iters
    order = blocks if op == "sa_layer_fwd" else blocks[::-1]
    for k in range(per):
        ms = sum(ev[i * per + k][0].elapsed_time(ev[i * per + k][1]) for i in range(iters)) / iters
        by, fl = ev[k][2]
        blk = order[k // 3] if per == 15 else "?"
        layer = (k % 3) if op == "sa_layer_fwd" else 2 - (k % 3)
        peak = bf16 if op == "sa_layer_bwd" else bf16 / 2
        total[op] += ms
        flops[op] += fl
        print("%-4s L%d %-13s %8.1f us  %7.2f GFLOP  %6.1f TFLOP/s = %.3f of the %s tensor peak   "
              "%7.1f MB block-level  %6.1f GB/s = %.3f of HBM"
              % (blk, layer, op, ms * 1e3, fl / 1e9, fl / ms / 1e9, fl / ms / 1e9 / peak,
                 "bf16" if op == "sa_layer_bwd" else "tf32", by / 1e6, by / ms / 1e6, by / ms / 1e6 / hbm))
for op in total:
    peak = bf16 if op == "sa_layer_bwd" else bf16 / 2
    print("%s: %.3f ms per step, %.1f GFLOP -> %.1f TFLOP/s = %.3f of the tensor peak"
          % (op, total[op], flops[op] / 1e9, flops[op] / total[op] / 1e9, flops[op] / total[op] / 1e9 / peak))
print("sum of fused SA-layer launches per step: %.3f ms" % sum(total.values()))
