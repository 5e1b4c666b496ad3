"""CUDA-event timing of every fused SA-layer launch at the backbone's four SA shapes (B=8):
python scripts/time_sa.py [iters] -> per-launch microseconds, algorithmic GB/s, fraction of the
measured HBM peak.  Inputs rotate over buffers larger than L2 between iterations."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backtoreality_b200 import _ext, scenes  # noqa: E402
from backtoreality_b200.pointnet2_modules import PointnetSAModuleVotes  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda:0")
torch.manual_seed(0)
B = 8
peak = 6553.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
pc = torch.from_numpy(scenes.batch(1000, B, 40000, C=1, kind="room", dup=0.2)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SHAPES = [("sa1", 40000, 1, dict(npoint=2048, radius=0.2, nsample=64, mlp=[1, 64, 64, 128])),
          ("sa2", 2048, 128, dict(npoint=1024, radius=0.4, nsample=32, mlp=[128, 128, 128, 256])),
          ("sa3", 1024, 256, dict(npoint=512, radius=0.8, nsample=16, mlp=[256, 128, 128, 256])),
          ("sa4", 512, 256, dict(npoint=256, radius=1.2, nsample=16, mlp=[256, 128, 128, 256]))]
_ext.TIME_OPS.update(["sa_layer_fwd", "sa_layer_bwd"])
total_ms = 0.0
for name, N, C, kw in SHAPES:
    sa = PointnetSAModuleVotes(use_xyz=True, normalize_xyz=True, **kw).to(dev).train()
    xyz = pc[:, :N, :3].contiguous()
    feats = (pc[..., 3:].transpose(1, 2).contiguous() if name == "sa1"
             else torch.randn(B, C, N, device=dev).requires_grad_(True))
    for it in range(iters + 2):
        if it == 2:
            torch.cuda.synchronize()
            _ext.TIMED.clear()
        flush.zero_()
        new_xyz, y, inds = sa(xyz, feats)
        y.square().mean().backward()
    torch.cuda.synchronize()
    for op in ("sa_layer_fwd", "sa_layer_bwd"):
        ev = _ext.TIMED.get(op, [])
        per = len(ev) // iters
        for k in range(per):
            ms = sum(ev[i * per + k][0].elapsed_time(ev[i * per + k][1]) for i in range(iters)) / iters
            by = ev[k][2]
            total_ms += ms
            print("%-4s %-13s launch %d  %8.1f us  %8.1f MB algorithmic  %7.1f GB/s  %.3f of measured HBM peak"
                  % (name, op, k, ms * 1e3, by / 1e6, by / ms / 1e6, by / ms / 1e6 / peak))
    _ext.TIMED.clear()
print("sum of fused SA-layer launches per step: %.3f ms" % total_ms)
