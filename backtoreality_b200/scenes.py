"""Synthetic ScanNet-shaped scenes -- the input contract of the hot path.

The reference feeds `(N, 3+C)` float32 point clouds, N = `num_points`, drawn from a scan with
`random_sampling` WITH replacement when the scan is smaller than N
(/root/reference/detection/Votenet/utils/pc_util.py:36-44), so exact duplicate points are
normal; C = 1 is the height feature `z - percentile(z, 0.99)`
(/root/reference/detection/Votenet/scannet/scannet_detection_dataset.py:122-125).

    scene(seed, N, C, kind)          one (N, 3+C) float32 array
    batch(first_index, B, N, C, ...) (B, N, 3+C), scene i uses seed 1000 + first_index + i

kinds: "uniform" = rand(N,3+C) in [0,1) (the reference's own smoke input,
models/backbone_module.py:374); "room" = axis-aligned room with floor / walls / box furniture
surfaces, ~200 pts/m^2, origin at the floor centre (so the |p|^2 <= 1e-3 hole of FPS is
exercised); "room_shifted" = the same moved by +5 m.
"""
import numpy as np


def _room_points(rng, n):
    lx, ly = rng.uniform(3.0, 9.0, size=2)
    lz = rng.uniform(2.4, 3.2)
    n_floor = int(0.35 * n)
    n_wall = int(0.35 * n)
    n_box = n - n_floor - n_wall
    pts = []
    # floor
    f = np.empty((n_floor, 3))
    f[:, 0] = rng.uniform(-lx / 2, lx / 2, n_floor)
    f[:, 1] = rng.uniform(-ly / 2, ly / 2, n_floor)
    f[:, 2] = 0.0
    pts.append(f)
    # four walls, points spread by wall area
    perim = 2 * (lx + ly)
    u = rng.uniform(0.0, perim, n_wall)
    w = np.empty((n_wall, 3))
    w[:, 2] = rng.uniform(0.0, lz, n_wall)
    a = u < lx
    b = (u >= lx) & (u < lx + ly)
    c = (u >= lx + ly) & (u < 2 * lx + ly)
    d = u >= 2 * lx + ly
    w[a, 0] = u[a] - lx / 2;            w[a, 1] = -ly / 2
    w[b, 0] = lx / 2;                   w[b, 1] = u[b] - lx - ly / 2
    w[c, 0] = u[c] - lx - ly - lx / 2;  w[c, 1] = ly / 2
    w[d, 0] = -lx / 2;                  w[d, 1] = u[d] - 2 * lx - ly - ly / 2
    pts.append(w)
    # furniture: surfaces of random boxes standing on the floor
    nb = int(rng.integers(5, 26))
    size = rng.uniform(0.3, 2.0, size=(nb, 3))
    size[:, 2] = np.minimum(size[:, 2], lz * 0.8)
    ctr = np.stack([rng.uniform(-lx / 2 + 0.2, lx / 2 - 0.2, nb),
                    rng.uniform(-ly / 2 + 0.2, ly / 2 - 0.2, nb)], axis=1)
    area = 2 * (size[:, 0] * size[:, 2] + size[:, 1] * size[:, 2]) + size[:, 0] * size[:, 1]
    which = rng.choice(nb, size=n_box, p=area / area.sum())
    q = rng.uniform(-0.5, 0.5, size=(n_box, 3))
    face = rng.integers(0, 5, n_box)  # 4 sides + top
    q[face == 0, 0] = -0.5
    q[face == 1, 0] = 0.5
    q[face == 2, 1] = -0.5
    q[face == 3, 1] = 0.5
    q[face == 4, 2] = 0.5
    bx = np.empty((n_box, 3))
    bx[:, 0] = ctr[which, 0] + q[:, 0] * size[which, 0]
    bx[:, 1] = ctr[which, 1] + q[:, 1] * size[which, 1]
    bx[:, 2] = (q[:, 2] + 0.5) * size[which, 2]
    pts.append(bx)
    p = np.concatenate(pts, axis=0)
    p += rng.normal(0.0, 0.005, size=p.shape)  # 5 mm sensor noise
    return p


def scene(seed, N, C=1, kind="room", dup=0.2, augment=True):
    """One synthetic scene: (N, 3+C) float32.  `dup` = fraction of exact duplicates produced by
    sampling with replacement from (1-dup)*N unique points."""
    rng = np.random.Generator(np.random.PCG64(int(seed)))
    if kind == "uniform":
        return rng.random((N, 3 + C), dtype=np.float32)
    if kind not in ("room", "room_shifted"):
        raise ValueError("unknown scene kind %r" % (kind,))
    n_unique = max(1, int(round((1.0 - dup) * N)))
    p = _room_points(rng, n_unique)
    if augment:  # the dataset's flips and +-5 degree z rotation (scannet_detection_dataset.py:147-162)
        if rng.random() > 0.5:
            p[:, 0] = -p[:, 0]
        if rng.random() > 0.5:
            p[:, 1] = -p[:, 1]
        ang = rng.uniform(-np.pi / 36, np.pi / 36)
        c, s = np.cos(ang), np.sin(ang)
        p[:, :2] = p[:, :2] @ np.array([[c, -s], [s, c]]).T
    if kind == "room_shifted":
        p[:, :2] += 5.0
    if n_unique < N:
        sel = np.concatenate([np.arange(n_unique), rng.integers(0, n_unique, N - n_unique)])
        sel = sel[rng.permutation(N)]
        p = p[sel]
    out = np.empty((N, 3 + C), dtype=np.float32)
    out[:, :3] = p.astype(np.float32)
    if C >= 1:
        floor = np.percentile(out[:, 2], 0.99)
        out[:, 3] = out[:, 2] - floor
    if C > 1:
        out[:, 4:] = rng.random((N, C - 1), dtype=np.float32)
    return out


def batch(first_index, B, N, C=1, kind="room", dup=0.2):
    """(B, N, 3+C) float32; scene i uses seed 1000 + first_index + i."""
    return np.stack([scene(1000 + first_index + i, N, C, kind, dup) for i in range(B)], axis=0)
