"""How many spatial buckets does an FPS iteration touch?  numpy simulation behind csrc/fps_bucket.cu:
points of one room scene sorted by a 30-bit Morton code, buckets of 64..1280 consecutive points with
their bounding boxes; per iteration a bucket is touched iff the squared distance from the new sample to
its box is below the bucket's largest running min-distance.  (40000 points, 2048 samples: ~8 of 63
buckets of 640 points per iteration, ~13 of 313 buckets of 128.)"""
import numpy as np, sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from backtoreality_b200 import scenes
N = 40000; npoint = 2048
pc = scenes.batch(1000, 1, N, C=1, kind="room", dup=0.2)[0]
xyz = pc[:, :3].astype(np.float32)
lo, hi = xyz.min(0), xyz.max(0)
q = np.clip(((xyz - lo) / (hi - lo + 1e-9) * 1023).astype(np.int64), 0, 1023)
def part(v):
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v
code = part(q[:, 0]) | (part(q[:, 1]) << 1) | (part(q[:, 2]) << 2)
order = np.argsort(code, kind="stable")
xs = xyz[order]
for bucket in (64, 128, 256, 640, 1280):
    nb = (N + bucket - 1) // bucket
    bid = np.arange(N) // bucket
    blo = np.stack([np.minimum.reduceat(xs[:, d], np.arange(0, N, bucket)) for d in range(3)], 1)
    bhi = np.stack([np.maximum.reduceat(xs[:, d], np.arange(0, N, bucket)) for d in range(3)], 1)
    temp = np.full(N, 1e10, np.float32)
    old = np.where(order == 0)[0][0]
    aff_hist = []
    bmax = np.full(nb, 1e10, np.float32)
    for it in range(npoint - 1):
        o = xs[old]
        d = np.maximum(np.maximum(blo - o, o - bhi), 0)
        lb = (d * d).sum(1)
        aff = lb < bmax
        aff_hist.append(aff.sum())
        dd = ((xs - o) ** 2).sum(1)
        temp = np.minimum(temp, dd)
        bmax = np.maximum.reduceat(temp, np.arange(0, N, bucket))
        old = int(np.argmax(temp))
    a = np.array(aff_hist)
    print("bucket %5d: %4d buckets; affected per iteration: mean %.1f (%.1f%%), median %d, last-1000 mean %.1f, max %d" % (
        bucket, nb, a.mean(), 100 * a.mean() / nb, np.median(a), a[-1000:].mean(), a.max()))
