"""Which aten ops launch the small copy / elementwise kernels of one training step?"""
import os, sys
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backtoreality_b200 import scenes
from backtoreality_b200.votenet import VoteNet
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = VoteNet(22, 1, 22, np.ones((22, 3), np.float32), input_feature_dim=1, num_proposal=256,
              vote_factor=1, sampling="vote_fps").to(dev).train()
pc = torch.from_numpy(scenes.batch(0, 8, 40000, C=1, kind="room", dup=0.2)).to(dev)
def step():
    for p in net.parameters(): p.grad = None
    ep = net({"point_clouds": pc})
    loss = (ep["proposal_scores_raw"] ** 2).mean() + ((ep["vote_xyz"] - ep["seed_xyz"]) ** 2).mean()
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True)
        if e.key in ("aten::copy_", "aten::contiguous", "aten::clone", "aten::cat", "aten::add_", "aten::add",
                     "aten::mul", "aten::div", "aten::sum", "aten::threshold_backward", "aten::relu_", "aten::fill_",
                     "aten::zero_", "aten::_to_copy", "aten::transpose")]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:40]:
    print("%-26s n=%3d cuda %8.1f us  %s" % (e.key, e.count, e.device_time_total, str(e.input_shapes)[:110]))
