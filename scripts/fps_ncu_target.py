"""Target of the FPS ncu capture (scripts/gpu_r2_profiles.sh): SA1-shaped FPS (8 scenes x 40000
points -> 2048) at the standalone cluster width and at the 4-CTA width used beside the step, then
the 2048 -> 1024 level."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backtoreality_b200 import _ext, scenes
dev = torch.device("cuda:0")
xyz = torch.from_numpy(scenes.batch(1000, 8, 40000, C=0, kind="room", dup=0.2)).to(dev)[..., :3].contiguous()
for _ in range(2):
    a = _ext.furthest_point_sampling(xyz, 2048)
    b = _ext.furthest_point_sampling(xyz, 2048, cluster=4)
    small = torch.gather(xyz, 1, a.long()[..., None].expand(-1, -1, 3)).contiguous()
    c = _ext.furthest_point_sampling(small, 1024)
torch.cuda.synchronize()
print("done", bool(torch.equal(a, b)))
