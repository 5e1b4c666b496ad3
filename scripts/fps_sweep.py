"""FPS timing: bucket kernel (csrc/fps_bucket.cu) vs round 1's kernel (csrc/fps.cu) over cluster
widths, B=8 room scenes, plus the later levels' sizes.  CUDA events, 5 launches after 2 warm-ups.
The sort pre-pass of the bucket kernel is inside its time."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backtoreality_b200 import _ext, scenes  # noqa: E402

dev = torch.device("cuda:0")
B = 8


def time_fps(xyz, npnt, cluster, legacy):
    _ext.FPS_LEGACY = legacy
    for _ in range(2):
        out = _ext.furthest_point_sampling(xyz, npnt, cluster=cluster)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = _ext.furthest_point_sampling(xyz, npnt, cluster=cluster)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5, out


for N, npnt in ((40000, 2048), (50000, 2048), (20000, 2048)):
    pc = torch.from_numpy(scenes.batch(1000, B, N, C=0, kind="room", dup=0.2)).to(dev)
    xyz = pc[..., :3].contiguous()
    ref = None
    for legacy in (True, False):
        for c in (0, 2, 3, 4, 5, 6, 8, 10, 12, 16):
            try:
                ms, out = time_fps(xyz, npnt, c, legacy)
            except Exception as e:
                print("N=%d cluster=%s legacy=%s failed: %s" % (N, c, legacy, str(e)[:100]))
                continue
            if ref is None:
                ref = out.clone()
            print("N=%d np=%d %-6s cluster=%-4s  %.3f ms  (%.0f ns/iteration, %.4f ms/scene)  same=%s"
                  % (N, npnt, "fps.cu" if legacy else "bucket", c or "auto", ms, ms * 1e6 / (npnt - 1),
                     ms / B, bool(torch.equal(out, ref))), flush=True)
cur = xyz
for N, npnt in ((2048, 1024), (1024, 512), (512, 256), (1024, 256)):
    pts = torch.from_numpy(scenes.batch(7, B, N, C=0, kind="room", dup=0.0)).to(dev)[..., :3].contiguous()
    for legacy in (True, False):
        ms, out = time_fps(pts, npnt, 0, legacy)
        print("N=%d np=%d %-6s  %.3f ms  (%.0f ns/iteration)" % (N, npnt, "fps.cu" if legacy else "bucket",
                                                                 ms, ms * 1e6 / (npnt - 1)), flush=True)
