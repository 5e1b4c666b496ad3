"""FPS kernel timing vs cluster size (B2R_FPS_CLUSTER override), B=8 room scenes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backtoreality_b200 import _ext, scenes  # noqa: E402

dev = torch.device("cuda:0")
B = 8
for N, npnt in ((40000, 2048), (50000, 2048)):
    pc = torch.from_numpy(scenes.batch(1000, B, N, C=0, kind="room", dup=0.2)).to(dev)
    xyz = pc[..., :3].contiguous()
    ref = None
    for c in (0, 6, 8, 10, 12, 14, 16):
        os.environ["B2R_FPS_CLUSTER"] = str(c)
        for _ in range(2):
            out = _ext.furthest_point_sampling(xyz, npnt)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = _ext.furthest_point_sampling(xyz, npnt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        if ref is None:
            ref = out.clone()
        print("N=%d np=%d cluster=%s  %.3f ms  (%.0f ns/iteration)  same=%s"
              % (N, npnt, c or "auto", ms, ms * 1e6 / (npnt - 1), bool(torch.equal(out, ref))))
