"""Run the SA1- (or SA2-) shaped fused block forward+backward a few times: the target of the
`ncu --set full` captures under profiles/ (python scripts/profile_sa.py [sa1|sa2] [iters])."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backtoreality_b200 import scenes  # noqa: E402
from backtoreality_b200.pointnet2_modules import PointnetSAModuleVotes  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "sa1"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
torch.manual_seed(0)
B = 8
if which == "all":
    # the five SA blocks of the VoteNet step (sa1..sa4 + vote aggregation), B = 8, one after the
    # other: 15 forward + 15 backward fused-layer launches per iteration, as in bench.py
    pc = torch.from_numpy(scenes.batch(1000, B, 40000, C=1, kind="room", dup=0.2)).to(dev)
    cfgs = [(40000, 1, dict(npoint=2048, radius=0.2, nsample=64, mlp=[1, 64, 64, 128])),
            (2048, 128, dict(npoint=1024, radius=0.4, nsample=32, mlp=[128, 128, 128, 256])),
            (1024, 256, dict(npoint=512, radius=0.8, nsample=16, mlp=[256, 128, 128, 256])),
            (512, 256, dict(npoint=256, radius=1.2, nsample=16, mlp=[256, 128, 128, 256])),
            (1024, 256, dict(npoint=256, radius=0.3, nsample=16, mlp=[256, 128, 128, 128]))]
    mods = [PointnetSAModuleVotes(use_xyz=True, normalize_xyz=True, **kw).to(dev).train() for _, _, kw in cfgs]
    for it in range(iters):
        for (N, C, kw), sa in zip(cfgs, mods):
            xyz = pc[:, :N, :3].contiguous()
            feats = (pc[..., 3:].transpose(1, 2).contiguous() if C == 1
                     else torch.randn(B, C, N, device=dev).requires_grad_(True))
            new_xyz, y, inds = sa(xyz, feats)
            y.square().mean().backward()
    torch.cuda.synchronize()
    print("done all", iters)
    sys.exit(0)
if which == "sa1":
    sa = PointnetSAModuleVotes(npoint=2048, radius=0.2, nsample=64, mlp=[1, 64, 64, 128],
                               use_xyz=True, normalize_xyz=True).to(dev).train()
    pc = torch.from_numpy(scenes.batch(1000, B, 40000, C=1, kind="room", dup=0.2)).to(dev)
    xyz = pc[..., :3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous()
else:
    sa = PointnetSAModuleVotes(npoint=1024, radius=0.4, nsample=32, mlp=[128, 128, 128, 256],
                               use_xyz=True, normalize_xyz=True).to(dev).train()
    pc = torch.from_numpy(scenes.batch(1000, B, 40000, C=1, kind="room", dup=0.2)).to(dev)
    xyz = pc[:, :2048, :3].contiguous()
    feats = torch.randn(B, 128, 2048, device=dev).requires_grad_(True)
for it in range(iters):
    new_xyz, y, inds = sa(xyz, feats)
    y.square().mean().backward()
torch.cuda.synchronize()
print("done", which, iters)
