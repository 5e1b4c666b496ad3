"""Per-op timing of libb2r vs the reference's own CUDA kernels (oracle/_ref/_ext.so, built for
sm_100a from the reference sources) on identical synthetic scenes.  CUDA events, warm-up 5,
median of `iters`.  Usage: python scripts/microbench.py [--B 8] [--N 40000] [--json out.json]"""
import argparse
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backtoreality_b200 import _ext as ours  # noqa: E402
from backtoreality_b200 import scenes  # noqa: E402


def load_ref():
    p = os.path.join(ROOT, "oracle", "_ref", "_ext.so")
    if not os.path.isfile(p):
        return None
    spec = importlib.util.spec_from_file_location("_ext", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--N", type=int, default=40000)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    ref = load_ref()
    pc = torch.from_numpy(scenes.batch(0, a.B, a.N, C=1, kind="room", dup=0.2)).to(dev)
    xyz = pc[..., :3].contiguous()
    res = {"B": a.B, "N": a.N}

    def both(name, f_ours, f_ref, same=None):
        t_o = timeit(f_ours)
        t_r = timeit(f_ref) if ref is not None else None
        ok = None
        if ref is not None and same is not None:
            ok = bool(same(f_ours(), f_ref()))
        res[name] = {"ours_ms": t_o, "ref_ms": t_r, "match": ok}
        print("%-28s ours %8.3f ms   ref %s   match=%s" % (
            name, t_o, "%8.3f ms" % t_r if t_r is not None else "   n/a", ok), flush=True)

    eq = lambda x, y: torch.equal(x, y)
    levels = [(a.N, 2048, 0.2, 64), (2048, 1024, 0.4, 32), (1024, 512, 0.8, 16), (512, 256, 1.2, 16)]
    cur = xyz
    for (n, npnt, r, ns) in levels:
        both("fps N=%d np=%d" % (n, npnt),
             lambda: ours.furthest_point_sampling(cur, npnt),
             lambda: ref.furthest_point_sampling(cur, npnt), eq)
        inds = ours.furthest_point_sampling(cur, npnt)
        new = torch.gather(cur, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
        both("ball_query N=%d np=%d ns=%d" % (n, npnt, ns),
             lambda: ours.ball_query(new, cur, r, ns),
             lambda: ref.ball_query(new, cur, r, ns), eq)
        idx = ours.ball_query(new, cur, r, ns)
        C = {a.N: 1, 2048: 128, 1024: 256, 512: 256}[n]
        f = torch.randn(a.B, C, n, device=dev)
        both("group_fwd C=%d" % C, lambda: ours.group_points(f, idx),
             lambda: ref.group_points(f, idx), eq)
        g = torch.randn(a.B, C, npnt, ns, device=dev)
        both("group_bwd C=%d" % C, lambda: ours.group_points_grad(g, idx, n),
             lambda: ref.group_points_grad(g, idx, n),
             lambda x, y: torch.allclose(x, y, rtol=1e-4, atol=1e-4))
        cur = new
    unk = torch.rand(a.B, 1024, 3, device=dev)
    kn = torch.rand(a.B, 512, 3, device=dev)
    both("three_nn 1024x512", lambda: ours.three_nn(unk, kn)[1], lambda: ref.three_nn(unk, kn)[1], eq)
    d2, i3 = ours.three_nn(unk, kn)
    w = torch.rand(a.B, 1024, 3, device=dev)
    f = torch.randn(a.B, 256, 512, device=dev)
    both("interp_fwd C=256", lambda: ours.three_interpolate(f, i3, w),
         lambda: ref.three_interpolate(f, i3, w), eq)
    g = torch.randn(a.B, 256, 1024, device=dev)
    both("interp_bwd C=256", lambda: ours.three_interpolate_grad(g, i3, w, 512),
         lambda: ref.three_interpolate_grad(g, i3, w, 512),
         lambda x, y: torch.allclose(x, y, rtol=1e-4, atol=1e-4))
    if a.json:
        os.makedirs(os.path.dirname(os.path.abspath(a.json)), exist_ok=True)
        json.dump(res, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
