"""HBM roofline of the standalone index movers (BASELINE.json configs[4] sweep): gather, grouping,
three_interpolate forward and backward.  For every shape: median ms (CUDA events, 20 launches after
5 warm-ups, a 256 MB L2 flush between launches), algorithmic bytes (every input
and output element once, indices once), GB/s and the fraction of the measured HBM copy peak
(MEASURED_PEAKS.json hbm_gbs, else the profiling guide's 6553 GB/s).
Usage: python scripts/movers_roofline.py [--quick] [--json out.json]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backtoreality_b200 import _ext as ours  # noqa: E402
from backtoreality_b200 import scenes  # noqa: E402


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6553.0, "fallback (B200_PROFILING.md)"


_flush = None


def timeit(fn, iters=20, warm=5):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        _flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    peak, src = peak_gbs()
    rows = []

    def report(name, shape, ms, nbytes):
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append({"op": name, "shape": shape, "ms": ms, "MB": nbytes / 1e6, "GBs": gbs, "frac": gbs / peak})
        print("%-22s %-34s %8.3f ms %9.1f MB %8.1f GB/s  %.3f of HBM peak" % (name, shape, ms, nbytes / 1e6, gbs, gbs / peak),
              flush=True)

    print("HBM peak %.0f GB/s (%s)" % (peak, src))
    # (B, N, npoint, nsample, radius, C): the four SA levels of the detectors + the sweep's corners
    cases = [(8, 40000, 2048, 64, 0.2, 4), (8, 2048, 1024, 32, 0.4, 128), (8, 1024, 512, 16, 0.8, 256),
             (1, 40000, 2048, 64, 0.2, 64), (1, 4096, 256, 16, 0.4, 128), (8, 4096, 1024, 32, 0.3, 64),
             (8, 20000, 2048, 32, 0.2, 32), (1, 100000, 2048, 64, 0.2, 16), (8, 100000, 2048, 16, 0.2, 8)]
    if a.quick:
        cases = cases[:3]
    for (B, N, NP, NS, r, C) in cases:
        xyz = torch.from_numpy(scenes.batch(0, B, N, C=0, kind="room", dup=0.2)).to(dev)[..., :3].contiguous()
        inds = ours.furthest_point_sampling(xyz, NP)
        new = torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
        idx = ours.ball_query(new, xyz, r, NS)
        f = torch.randn(B, C, N, device=dev)
        shape = "B=%d N=%d np=%d ns=%d C=%d" % (B, N, NP, NS, C)
        report("gather_fwd", shape, timeit(lambda: ours.gather_points(f, inds)), 4 * B * (C * NP * 2 + NP))
        gg = torch.randn(B, C, NP, device=dev)
        report("gather_bwd", shape, timeit(lambda: ours.gather_points_grad(gg, inds, N)), 4 * B * (C * NP + NP + C * N))
        L = NP * NS
        # forward: every source row may be read (<= C*N), idx once, output once
        report("group_fwd", shape, timeit(lambda: ours.group_points(f, idx)), 4 * B * (C * L + L + C * N))
        g = torch.randn(B, C, NP, NS, device=dev)
        report("group_bwd", shape, timeit(lambda: ours.group_points_grad(g, idx, N)), 4 * B * (C * L + L + C * N))
        del f, g, gg
    # three_interpolate: (B, n unknown, m known, C): fp1 / fp2 of the detectors + sweep corners
    icases = [(8, 512, 256, 256), (8, 1024, 512, 256), (8, 2048, 1024, 256), (8, 40000, 2048, 128),
              (1, 40000, 2048, 128), (1, 100000, 4096, 64)]
    if a.quick:
        icases = icases[:2]
    for (B, n, m, C) in icases:
        unk = torch.rand(B, n, 3, device=dev)
        kn = torch.rand(B, m, 3, device=dev)
        d2, i3 = ours.three_nn(unk, kn)
        w = torch.rand(B, n, 3, device=dev)
        w = (w / w.sum(2, keepdim=True)).contiguous()
        f = torch.randn(B, C, m, device=dev)
        shape = "B=%d n=%d m=%d C=%d" % (B, n, m, C)
        report("three_nn", shape, timeit(lambda: ours.three_nn(unk, kn)), 4 * B * (3 * n + 3 * m + 6 * n))
        report("three_interp_fwd", shape, timeit(lambda: ours.three_interpolate(f, i3, w)), 4 * B * (C * m + 6 * n + C * n))
        g = torch.randn(B, C, n, device=dev)
        report("three_interp_bwd", shape, timeit(lambda: ours.three_interpolate_grad(g, i3, w, m)), 4 * B * (C * n + 6 * n + C * m))
    if a.json:
        os.makedirs(os.path.dirname(os.path.abspath(a.json)), exist_ok=True)
        json.dump({"hbm_peak_gbs": peak, "peak_source": src, "rows": rows}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
