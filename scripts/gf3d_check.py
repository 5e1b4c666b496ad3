"""BASELINE.json configs[3] shape check: GroupFree3D backbone (input_feature_dim 0, fp2 width 288),
B = 4 scenes of 50000 points, forward + backward through the fused path, eager and graph-replayed."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backtoreality_b200 import scenes  # noqa: E402
from backtoreality_b200.backbone_module import Pointnet2Backbone  # noqa: E402
from backtoreality_b200.train_step import CapturedTrainStep, PipelinedTrainStep  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = Pointnet2Backbone(input_feature_dim=0, fp2_out=288).to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True, capturable=True)
pcs = [torch.from_numpy(scenes.batch(500 + 4 * i, 4, 50000, C=0, kind="room", dup=0.2)).to(dev) for i in range(3)]


def step(pc, geometry=None):
    for p in net.parameters():
        p.grad = None
    ep = net(pc, geometry=geometry)
    loss = ep["fp2_features"].square().mean()
    loss.backward()
    opt.step()
    return loss.detach()


def time(fn, n=10):
    for i in range(3):
        fn(pcs[i % 3])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(pcs[i % 3])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


ms_e = time(step)
cap = CapturedTrainStep(step, pcs[0])
ms_g = time(lambda pc: cap(pc))
pipe = PipelinedTrainStep(net, step, pcs[0], fps_cluster=5)   # 50000 points: 20 per thread on 5 CTAs
ms_p = time(lambda pc: pipe(pc))
print("GF3D backbone fwd+bwd+Adam, B=4 x 50000 points: eager %.2f ms/step (%.0f scenes/s), "
      "graph %.2f ms/step (%.0f scenes/s), pipelined graph %.2f ms/step (%.0f scenes/s), loss %.4f"
      % (ms_e, 4e3 / ms_e, ms_g, 4e3 / ms_g, ms_p, 4e3 / ms_p, float(cap(pcs[0]))))
