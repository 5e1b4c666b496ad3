"""Run-to-run variability of gradients (same inputs, same weights)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import rel_l2
from backtoreality_b200 import scenes, fused_sa
from backtoreality_b200.votenet import VoteNet
cuda = torch.device("cuda:0")
torch.manual_seed(5)
net = VoteNet(4, 1, 4, np.ones((4, 3), np.float32), input_feature_dim=1, num_proposal=64,
              vote_factor=1, sampling="vote_fps").to(cuda).train()
pc = torch.from_numpy(scenes.batch(300, 2, 8192, C=1, kind="room", dup=0.2)).to(cuda)
for fused in (True, False):
    fused_sa.ENABLED = fused
    runs = []
    for r in range(3):
        for p in net.parameters():
            p.grad = None
        ep = net({"point_clouds": pc})
        loss = (ep["proposal_scores_raw"] ** 2).mean() + ((ep["vote_xyz"] - ep["seed_xyz"]) ** 2).mean()
        loss.backward()
        runs.append((float(loss), {n: p.grad.clone() for n, p in net.named_parameters()}))
    print("fused" if fused else "unfused", "losses", [r[0] for r in runs])
    worst = sorted(((rel_l2(runs[1][1][n].cpu().numpy(), runs[0][1][n].cpu().numpy()), n) for n in runs[0][1]), reverse=True)[:6]
    for e, n in worst:
        print("   %.3e  %s" % (e, n))
