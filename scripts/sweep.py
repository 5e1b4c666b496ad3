"""BASELINE.json configs[4]: FPS / ball query / three_nn + three_interpolate microbench sweep,
libb2r vs the reference's own kernels (oracle/_ref/_ext.so) on the same B200.  B = 8 room scenes
(20 % duplicate points), CUDA events, median of 7 after 2 warm-ups; indices compared bit for bit."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backtoreality_b200 import _ext as ours  # noqa: E402
from backtoreality_b200 import scenes  # noqa: E402

p = os.path.join(ROOT, "oracle", "_ref", "_ext.so")
ref = None
if os.path.isfile(p):
    spec = importlib.util.spec_from_file_location("_ext", p)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
dev = torch.device("cuda:0")
B = 8


def timeit(fn, iters=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def row(name, f_o, f_r, same):
    t_o = timeit(f_o)
    t_r = timeit(f_r, iters=3, warm=1) if ref is not None else float("nan")
    ok = bool(same(f_o(), f_r())) if ref is not None else None
    print("%-44s ours %8.3f ms   ref %9.3f ms   x%6.1f   match=%s" % (name, t_o, t_r, t_r / t_o, ok), flush=True)


eq = lambda x, y: torch.equal(x, y)
print("== FPS (B=%d)" % B)
for N in (4096, 8192, 16384, 20000, 40000, 50000, 65536, 100000):
    xyz = torch.from_numpy(scenes.batch(7, B, N, C=0, kind="room", dup=0.2)).to(dev)[..., :3].contiguous()
    for npnt in ((256, 2048) if N not in (40000, 50000) else (256, 512, 1024, 2048)):
        row("fps N=%d npoint=%d" % (N, npnt), lambda: ours.furthest_point_sampling(xyz, npnt),
            lambda: ref.furthest_point_sampling(xyz, npnt), eq)
print("== ball query (B=%d)" % B)
for N in (4096, 16384, 40000, 100000):
    xyz = torch.from_numpy(scenes.batch(7, B, N, C=0, kind="room", dup=0.2)).to(dev)[..., :3].contiguous()
    for npnt in (256, 2048):
        inds = ours.furthest_point_sampling(xyz, npnt)
        new = torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
        for r, ns in ((0.2, 16), (0.2, 64), (0.4, 32)):
            row("ball_query N=%d np=%d r=%.1f ns=%d" % (N, npnt, r, ns), lambda: ours.ball_query(new, xyz, r, ns),
                lambda: ref.ball_query(new, xyz, r, ns), eq)
print("== three_nn / three_interpolate (B=%d)" % B)
for n, m, C in ((512, 256, 256), (1024, 512, 256), (2048, 1024, 128), (40000, 2048, 128)):
    unk = torch.rand(B, n, 3, device=dev)
    kn = torch.rand(B, m, 3, device=dev)
    row("three_nn n=%d m=%d" % (n, m), lambda: ours.three_nn(unk, kn)[1], lambda: ref.three_nn(unk, kn)[1], eq)
    i3 = ours.three_nn(unk, kn)[1]
    w = torch.rand(B, n, 3, device=dev)
    f = torch.randn(B, C, m, device=dev)
    row("three_interpolate fwd C=%d n=%d m=%d" % (C, n, m), lambda: ours.three_interpolate(f, i3, w),
        lambda: ref.three_interpolate(f, i3, w), eq)
    g = torch.randn(B, C, n, device=dev)
    row("three_interpolate bwd C=%d n=%d m=%d" % (C, n, m), lambda: ours.three_interpolate_grad(g, i3, w, m),
        lambda: ref.three_interpolate_grad(g, i3, w, m), lambda x, y: torch.allclose(x, y, rtol=1e-4, atol=1e-4))
