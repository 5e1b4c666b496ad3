"""Parity of every libb2r op (through the Python shim -> C ABI) against the CPU oracle.

Indices (FPS, ball query, three_nn) must be BIT-EXACT; movers are exact copies; interpolation
and the scatter-add backward passes are compared with fp32 tolerances stated per test.
Edge cases follow SURVEY.md appendix C.
"""
import numpy as np
import pytest
import torch

from _util import rel_l2
from backtoreality_b200 import scenes
from oracle import cpu_ops

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


# ------------------------------------------------------------------ FPS ---------------------
def _fps_check(xyz, npoint, dev):
    from backtoreality_b200 import _ext
    got = _ext.furthest_point_sampling(_t(xyz, dev), npoint).cpu().numpy()
    want = cpu_ops.fps(xyz, npoint)
    bad = np.argwhere(got != want)
    assert bad.size == 0, "first mismatch at (b, j) = %s: got %d want %d" % (
        bad[0], got[tuple(bad[0])], want[tuple(bad[0])])


@pytest.mark.parametrize("N,npoint", [(20000, 2048), (40000, 2048), (50000, 2048)])
def test_fps_room_scenes_with_duplicates(cuda, N, npoint):
    xyz = scenes.batch(0, 2, N, C=0, kind="room", dup=0.2)[..., :3]
    _fps_check(xyz, npoint, cuda)


@pytest.mark.parametrize("N,npoint", [(2048, 1024), (1024, 512), (512, 256), (1024, 256)])
def test_fps_small_levels(cuda, N, npoint):
    xyz = scenes.batch(3, 4, N, C=0, kind="room_shifted", dup=0.2)[..., :3]
    _fps_check(xyz, npoint, cuda)


@pytest.mark.parametrize("N", [1, 2, 3, 5, 31, 32, 33, 100, 255, 256, 257, 511, 513, 700, 1000,
                               1023, 1025, 3000, 4096, 4097, 5000, 9999, 16384, 65536, 100000])
def test_fps_odd_sizes(cuda, N):
    rng = np.random.default_rng(N)
    xyz = rng.random((2, N, 3), dtype=np.float32) + 0.5
    _fps_check(xyz, min(N, 64), cuda)


def test_fps_npoint_equals_n_and_exceeds_distinct(cuda):
    rng = np.random.default_rng(7)
    base = rng.random((1, 40, 3), dtype=np.float32) + 1.0
    xyz = np.concatenate([base, base, base], axis=1)  # 120 points, 40 distinct: heavy ties
    _fps_check(xyz, 120, cuda)
    xyz = np.repeat(base, 20, axis=1)  # N=800 > 512: two rows per reference lane
    _fps_check(xyz, 300, cuda)


def test_fps_hole_and_invalid_points(cuda):
    rng = np.random.default_rng(11)
    xyz = (rng.random((3, 3000, 3), dtype=np.float32) - 0.5) * 4.0
    xyz[:, ::7] *= 0.004                       # many points with |p|^2 <= 1e-3
    xyz[0, 0] = 0.0                            # point 0 itself inside the hole
    s = np.float32(np.sqrt(1e-3 / 3.0))
    xyz[1, 5] = s                              # right at the boundary (decided in double)
    xyz[1, 6] = np.nextafter(s, np.float32(1)) # just outside
    _fps_check(xyz, 128, cuda)
    allbad = np.zeros((2, 600, 3), np.float32)  # every point invalid -> 0,0,0,...
    _fps_check(allbad, 16, cuda)
    same = np.ones((1, 777, 3), np.float32)     # all identical, valid
    _fps_check(same, 50, cuda)


@pytest.mark.parametrize("cluster", [1, 2, 4, 8, 16])
def test_fps_forced_cluster_sizes(cuda, cluster, monkeypatch):
    monkeypatch.setenv("B2R_FPS_CLUSTER", str(cluster))
    xyz = scenes.batch(5, 2, 12000, C=0, kind="room", dup=0.3)[..., :3]
    _fps_check(xyz, 300, cuda)


@pytest.mark.parametrize("hint", [1, 3, 4, 5, 10, 16])
def test_fps_cluster_hint_does_not_change_indices(cuda, hint):
    """b2r_fps_ex: the caller-chosen cluster width (narrow clusters for FPS that runs beside
    other kernels) changes the launch geometry only -- the indices stay the oracle's; a hint too
    narrow for the register-resident capacity is widened, not refused."""
    from backtoreality_b200 import _ext
    xyz = scenes.batch(9, 2, 40000, C=0, kind="room", dup=0.2)[..., :3]
    got = _ext.furthest_point_sampling(_t(xyz, cuda), 512, cluster=hint).cpu().numpy()
    assert np.array_equal(got, cpu_ops.fps(xyz, 512))
    small = scenes.batch(10, 2, 3000, C=0, kind="room", dup=0.2)[..., :3]  # one CTA regardless
    got = _ext.furthest_point_sampling(_t(small, cuda), 256, cluster=hint).cpu().numpy()
    assert np.array_equal(got, cpu_ops.fps(small, 256))


@pytest.mark.parametrize("N,npoint", [(40000, 1024), (3000, 256), (300, 64)])
def test_fps_legacy_kernel_still_exact(cuda, N, npoint, monkeypatch):
    """csrc/fps.cu (every point updated every iteration; b2r_fps / b2r_fps_ex) stays in the
    library as the workspace-free entry point: same indices as the bucket kernel and the oracle."""
    from backtoreality_b200 import _ext
    xyz = scenes.batch(21, 2, N, C=0, kind="room", dup=0.2)[..., :3]
    want = cpu_ops.fps(xyz, npoint)
    monkeypatch.setattr(_ext, "FPS_LEGACY", True)
    assert np.array_equal(_ext.furthest_point_sampling(_t(xyz, cuda), npoint).cpu().numpy(), want)
    monkeypatch.setattr(_ext, "FPS_LEGACY", False)
    assert np.array_equal(_ext.furthest_point_sampling(_t(xyz, cuda), npoint).cpu().numpy(), want)


def test_fps_bucket_kernel_degenerate_geometry(cuda):
    """Cases that stress the bucket bound: all points on a line / in one cell, coordinates far
    from the origin, a few far outliers (huge bounding box, every other point in one cell)."""
    rng = np.random.default_rng(31)
    line = np.zeros((2, 6000, 3), np.float32)
    line[..., 0] = rng.random((2, 6000), dtype=np.float32) * 50 + 1
    _fps_check(line, 200, cuda)
    far = rng.random((2, 9000, 3), dtype=np.float32) * 2 + np.float32(4000.0)
    _fps_check(far, 200, cuda)
    out = rng.random((2, 20000, 3), dtype=np.float32) + 1
    out[:, :5] *= 1e4
    _fps_check(out, 300, cuda)
    grid = np.stack(np.meshgrid(np.arange(20), np.arange(20), np.arange(20)), -1).reshape(1, -1, 3)
    _fps_check((grid.astype(np.float32) + 1.0), 400, cuda)     # massive exact ties (lattice)


def test_fps_npoint_zero_and_one(cuda):
    from backtoreality_b200 import _ext
    xyz = torch.rand(2, 100, 3, device=cuda)
    assert _ext.furthest_point_sampling(xyz, 0).shape == (2, 0)
    assert _ext.furthest_point_sampling(xyz, 1).cpu().tolist() == [[0], [0]]


# ------------------------------------------------------------------ ball query --------------
def _bq_check(new_xyz, xyz, r, ns, dev):
    from backtoreality_b200 import _ext
    got = _ext.ball_query(_t(new_xyz, dev), _t(xyz, dev), r, ns).cpu().numpy()
    want = cpu_ops.ball_query(new_xyz, xyz, float(np.float32(r)), ns)
    bad = np.argwhere(got != want)
    assert bad.size == 0, "first mismatch at %s: got %s want %s" % (
        bad[0], got[bad[0][0], bad[0][1]], want[bad[0][0], bad[0][1]])


@pytest.mark.parametrize("N,M,r,ns", [(20000, 2048, 0.2, 64), (2048, 1024, 0.4, 32),
                                      (1024, 512, 0.8, 16), (512, 256, 1.2, 16),
                                      (1024, 256, 0.3, 16)])
def test_ball_query_backbone_shapes(cuda, N, M, r, ns):
    xyz = scenes.batch(0, 2, N, C=0, kind="room", dup=0.2)[..., :3]
    inds = cpu_ops.fps(xyz, M)
    new_xyz = np.take_along_axis(xyz, inds[..., None].astype(np.int64), axis=1)
    _bq_check(new_xyz, xyz, r, ns, cuda)


def test_ball_query_dense_uniform_early_exit(cuda):
    rng = np.random.default_rng(3)
    xyz = rng.random((2, 40000, 3), dtype=np.float32)
    new_xyz = xyz[:, :500].copy()
    _bq_check(new_xyz, xyz, 0.2, 64, cuda)


@pytest.mark.parametrize("case", ["room", "negative", "cluster", "huge_radius", "far_aliasing",
                                  "nan", "ragged", "beyond_grid_range", "overflowing_coordinate"])
def test_ball_query_grid_path_matches_oracle(cuda, case, monkeypatch):
    """The hashed-grid ball query (b2r_ball_query_grid, taken for N >= 8192; forced here from
    N >= 1024) must return exactly what the brute-force scan returns: against the C oracle, on
    inputs that stress the grid -- negative coordinates, thousands of hits per centre (buffer
    compaction), scenes wider than the 64-cell hash period (aliasing), NaN points, N % 32 != 0."""
    from backtoreality_b200 import _ext
    monkeypatch.setattr(_ext, "BALL_QUERY_GRID_MIN_N", 0)
    rng = np.random.default_rng(11)
    if case == "room":
        xyz = scenes.batch(3, 2, 40000, C=0, kind="room", dup=0.2)[..., :3]
        new_xyz, r, ns = xyz[:, ::37][:, :1000].copy(), 0.2, 64
    elif case == "negative":
        xyz = (rng.random((2, 9000, 3), dtype=np.float32) - 0.5) * 6.0
        new_xyz, r, ns = xyz[:, :700].copy() + np.float32(0.01), 0.3, 32
    elif case == "cluster":      # 6000 identical points + a tight blob: > 512 hits per centre
        xyz = rng.random((2, 12000, 3), dtype=np.float32)
        xyz[:, 1000:7000] = np.float32([0.5, 0.5, 0.5])
        xyz[:, 7000:9000] = 0.5 + (rng.random((2, 2000, 3), dtype=np.float32) - 0.5) * 0.05
        new_xyz, r, ns = xyz[:, 900:1100].copy(), 0.2, 64
    elif case == "huge_radius":  # every centre sees a large share of the scene
        xyz = rng.random((2, 8192, 3), dtype=np.float32)
        new_xyz, r, ns = xyz[:, :64].copy(), 1.2, 16
    elif case == "far_aliasing":  # 40 m wide scene at r = 0.2: 200 cells > the 64-cell hash period
        xyz = rng.random((2, 20000, 3), dtype=np.float32) * np.float32([40.0, 40.0, 10.0])
        new_xyz, r, ns = xyz[:, :512].copy(), 0.2, 16
    elif case == "nan":
        xyz = rng.random((2, 5000, 3), dtype=np.float32)
        xyz[:, 17] = np.nan
        xyz[:, 4000, 1] = np.inf
        new_xyz = xyz[:, :300].copy()
        r, ns = 0.25, 32
    elif case == "beyond_grid_range":
        # a scan georeferenced 3 km from the origin at r = 0.2: 15000 cells, past the ~8000 the
        # grid's fp32 rounding margin covers -> the kernel falls back to the index-order scan
        xyz = rng.random((2, 6000, 3), dtype=np.float32) * np.float32(2.0) + np.float32(3000.0)
        xyz[1] -= np.float32(3000.0)            # scene 1 stays on the grid path
        new_xyz, r, ns = xyz[:, :400].copy(), 0.2, 16
    elif case == "overflowing_coordinate":
        # one finite coordinate whose cell index does not fit an int32 (x / cell ~ 5e37)
        xyz = rng.random((2, 3000, 3), dtype=np.float32)
        xyz[0, 5, 0] = np.float32(1e37)
        new_xyz, r, ns = xyz[:, 100:300].copy(), 0.25, 8
    else:
        xyz = rng.random((2, 1025, 3), dtype=np.float32)
        new_xyz, r, ns = rng.random((2, 9, 3), dtype=np.float32), 0.35, 5
    _bq_check(new_xyz, xyz, r, ns, cuda)


@pytest.mark.parametrize("N", [1, 3, 4, 5, 127, 128, 129, 1023, 1024, 1025, 2050])
def test_ball_query_ragged_sizes(cuda, N):
    rng = np.random.default_rng(N)
    xyz = rng.random((2, N, 3), dtype=np.float32)
    new_xyz = rng.random((2, 9, 3), dtype=np.float32)
    for ns in (1, 5, 16):
        _bq_check(new_xyz, xyz, 0.35, ns, cuda)


def test_ball_query_counts_0_1_ns_and_boundary(cuda):
    # a line of points spaced exactly 0.25 apart; r chosen so that d2 == r2 exactly occurs
    xs = np.arange(64, dtype=np.float32) * np.float32(0.25)
    xyz = np.stack([xs, np.zeros_like(xs), np.zeros_like(xs)], -1)[None]
    centres = np.array([[[0.0, 0.0, 0.0], [100.0, 0.0, 0.0], [8.0, 0.0, 0.0], [0.125, 0.0, 0.0]]],
                       np.float32)
    for r in (0.25, 0.5, 0.1, 1.0, 20.0):   # 0.25 and 0.5: points AT the radius are excluded
        for ns in (1, 4, 8):
            _bq_check(centres, xyz, r, ns, cuda)


# ------------------------------------------------------------------ three_nn ----------------
@pytest.mark.parametrize("n,m", [(512, 256), (1024, 512), (2048, 1024), (777, 1), (50, 2), (50, 3),
                                 (33, 1030), (5000, 2048)])
def test_three_nn_exact(cuda, n, m):
    from backtoreality_b200 import _ext
    rng = np.random.default_rng(n * 31 + m)
    unknown = rng.random((2, n, 3), dtype=np.float32)
    known = rng.random((2, m, 3), dtype=np.float32)
    if m >= 8:
        known[:, 4] = known[:, 1]      # equidistant neighbours: earlier index must win
        unknown[:, 0] = known[:, 2]    # unknown == known
    d2, idx = _ext.three_nn(_t(unknown, cuda), _t(known, cuda))
    wd2, widx = cpu_ops.three_nn(unknown, known)
    assert np.array_equal(idx.cpu().numpy(), widx)
    assert np.array_equal(d2.cpu().numpy(), wd2)      # bit-exact incl. +inf for m < 3


# ------------------------------------------------------------------ movers ------------------
@pytest.mark.parametrize("C,N,M", [(3, 40000, 2048), (288, 1024, 256), (1, 7, 5), (5, 100, 100)])
def test_gather_fwd_bwd(cuda, C, N, M):
    from backtoreality_b200 import _ext
    rng = np.random.default_rng(C + N + M)
    f = rng.standard_normal((2, C, N)).astype(np.float32)
    idx = rng.integers(0, N, (2, M)).astype(np.int32)
    idx[:, : M // 2] = idx[:, M // 2: M // 2 * 2]   # repeated targets for the scatter-add
    out = _ext.gather_points(_t(f, cuda), _t(idx, cuda)).cpu().numpy()
    assert np.array_equal(out, cpu_ops.gather(f, idx))
    g = rng.standard_normal((2, C, M)).astype(np.float32)
    gf = _ext.gather_points_grad(_t(g, cuda), _t(idx, cuda), N).cpu().numpy()
    np.testing.assert_allclose(gf, cpu_ops.gather_grad(g, idx, N), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("C,N,NP,NS", [(128, 2048, 1024, 32), (3, 20000, 2048, 64), (259, 1024, 256, 16),
                                       (2, 50, 7, 3), (1, 10, 1, 1)])
def test_group_fwd_bwd(cuda, C, N, NP, NS):
    from backtoreality_b200 import _ext
    rng = np.random.default_rng(C + N + NP)
    f = rng.standard_normal((2, C, N)).astype(np.float32)
    idx = rng.integers(0, N, (2, NP, NS)).astype(np.int32)
    out = _ext.group_points(_t(f, cuda), _t(idx, cuda)).cpu().numpy()
    assert np.array_equal(out, cpu_ops.group(f, idx))
    g = rng.standard_normal((2, C, NP, NS)).astype(np.float32)
    gf = _ext.group_points_grad(_t(g, cuda), _t(idx, cuda), N).cpu().numpy()
    # sums of up to NP*NS/N (+ hot spots) fp32 terms in arbitrary order
    np.testing.assert_allclose(gf, cpu_ops.group_grad(g, idx, N), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("C,m,n", [(256, 256, 512), (256, 512, 1024), (128, 2048, 40000), (3, 5, 7)])
def test_three_interpolate_fwd_bwd(cuda, C, m, n):
    from backtoreality_b200 import _ext
    rng = np.random.default_rng(C + m + n)
    f = rng.standard_normal((2, C, m)).astype(np.float32)
    idx = rng.integers(0, m, (2, n, 3)).astype(np.int32)
    w = rng.random((2, n, 3), dtype=np.float32)
    w /= w.sum(-1, keepdims=True)
    out = _ext.three_interpolate(_t(f, cuda), _t(idx, cuda), _t(w, cuda)).cpu().numpy()
    assert np.array_equal(out, cpu_ops.interp(f, idx, w))   # same contraction => bit-exact
    g = rng.standard_normal((2, C, n)).astype(np.float32)
    gf = _ext.three_interpolate_grad(_t(g, cuda), _t(idx, cuda), _t(w, cuda), m).cpu().numpy()
    np.testing.assert_allclose(gf, cpu_ops.interp_grad(g, idx, w, m), rtol=1e-4, atol=1e-4)


def test_scatter_backward_is_deterministic_and_matches_atomic_path(cuda, monkeypatch):
    """group_points_grad / three_interpolate_grad through the scatter plan (csrc/movers_staged.cu):
    bit-identical over repeated runs (no atomics), equal to the oracle and to round 1's atomic
    kernels within summation-order tolerance; shapes with several source tiles (NP*NS > 16384),
    hot targets (every ball pads with index 0), a row length that is not a multiple of 4."""
    from backtoreality_b200 import _ext
    monkeypatch.setattr(_ext, "DETERMINISTIC", True)   # plan path for every shape (B2R_DETERMINISTIC=1)
    rng = np.random.default_rng(77)
    for (C, N, NP, NS) in [(5, 40000, 2048, 64), (64, 2048, 1024, 32), (3, 301, 37, 5)]:
        idx = rng.integers(0, N, (2, NP, NS)).astype(np.int32)
        idx[:, :, NS // 2:] = idx[:, :, :1]          # padded balls: heavy duplicates
        idx[0, : NP // 2, 0] = 0                        # one hot target
        g = rng.standard_normal((2, C, NP, NS)).astype(np.float32)
        gt, it = _t(g, cuda), _t(idx, cuda)
        runs = [_ext.group_points_grad(gt, it, N) for _ in range(3)]
        assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2])
        np.testing.assert_allclose(runs[0].cpu().numpy(), cpu_ops.group_grad(g, idx, N), rtol=2e-4, atol=2e-4)
        monkeypatch.setattr(_ext, "SCATTER_ATOMIC", True)
        atomic = _ext.group_points_grad(gt, it, N)
        monkeypatch.setattr(_ext, "SCATTER_ATOMIC", False)
        torch.testing.assert_close(runs[0], atomic, rtol=2e-4, atol=2e-4)
    for (C, m, n) in [(16, 2048, 40000), (7, 33, 10001)]:
        idx = rng.integers(0, m, (2, n, 3)).astype(np.int32)
        w = rng.random((2, n, 3), dtype=np.float32)
        g = rng.standard_normal((2, C, n)).astype(np.float32)
        gt, it, wt = _t(g, cuda), _t(idx, cuda), _t(w, cuda)
        runs = [_ext.three_interpolate_grad(gt, it, wt, m) for _ in range(3)]
        assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2])
        np.testing.assert_allclose(runs[0].cpu().numpy(), cpu_ops.interp_grad(g, idx, w, m), rtol=2e-4, atol=3e-4)
        monkeypatch.setattr(_ext, "SCATTER_ATOMIC", True)
        atomic = _ext.three_interpolate_grad(gt, it, wt, m)
        monkeypatch.setattr(_ext, "SCATTER_ATOMIC", False)
        torch.testing.assert_close(runs[0], atomic, rtol=2e-4, atol=3e-4)


def test_scatter_plan_cache_follows_in_place_index_updates(cuda):
    """the plan cache is keyed on (storage, version): an index tensor modified in place gets a new plan"""
    from backtoreality_b200 import _ext
    rng = np.random.default_rng(5)
    idx = _t(rng.integers(0, 64, (1, 16, 8)).astype(np.int32), cuda)
    g = _t(rng.standard_normal((1, 4, 16, 8)).astype(np.float32), cuda)
    a = _ext.group_points_grad(g, idx, 64)
    idx.copy_(_t(rng.integers(0, 64, (1, 16, 8)).astype(np.int32), cuda))
    b = _ext.group_points_grad(g, idx, 64)
    np.testing.assert_allclose(b.cpu().numpy(), cpu_ops.group_grad(g.cpu().numpy(), idx.cpu().numpy(), 64),
                               rtol=1e-4, atol=1e-4)
    assert not torch.equal(a, b)


# ------------------------------------------------------------------ fused QueryAndGroup -----
@pytest.mark.parametrize("C,N,NP,NS,norm", [(1, 5000, 512, 64, True), (128, 2048, 1024, 32, True),
                                            (0, 3000, 100, 16, True), (4, 300, 13, 3, False)])
def test_query_group_matches_unfused_composition(cuda, C, N, NP, NS, norm):
    from backtoreality_b200 import _ext
    rng = np.random.default_rng(C + N)
    xyz = rng.random((2, N, 3), dtype=np.float32)
    new_xyz = rng.random((2, NP, 3), dtype=np.float32)
    f = rng.standard_normal((2, C, N)).astype(np.float32) if C else None
    idx = rng.integers(0, N, (2, NP, NS)).astype(np.int32)
    r = 0.3
    out = _ext.query_group(_t(xyz, cuda), _t(new_xyz, cuda), _t(f, cuda) if C else None,
                           _t(idx, cuda), r, norm).cpu().numpy()
    # the reference's composition (pointnet2_utils.py:347-359) on the oracle
    gx = cpu_ops.group(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx)
    gx = gx - new_xyz.transpose(0, 2, 1)[..., None]
    if norm:
        gx = gx / np.float32(r)
    want = np.concatenate([gx, cpu_ops.group(f, idx)], 1) if C else gx
    assert np.array_equal(out, want.astype(np.float32))

    g = rng.standard_normal(out.shape).astype(np.float32)
    gxyz, gnew, gf = _ext.query_group_grad(_t(g, cuda), _t(idx, cuda), N, C, r, norm,
                                           True, True, True)
    gg = g[:, :3] / np.float32(r) if norm else g[:, :3]
    want_gxyz = cpu_ops.group_grad(np.ascontiguousarray(gg), idx, N).transpose(0, 2, 1)
    want_gnew = -gg.sum(-1).transpose(0, 2, 1)
    np.testing.assert_allclose(gxyz.cpu().numpy(), want_gxyz, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(gnew.cpu().numpy(), want_gnew, rtol=1e-4, atol=1e-4)
    if C:
        want_gf = cpu_ops.group_grad(np.ascontiguousarray(g[:, 3:]), idx, N)
        np.testing.assert_allclose(gf.cpu().numpy(), want_gf, rtol=1e-4, atol=1e-4)
    else:
        assert gf is None


# ------------------------------------------------------------------ boundary contract -------
def test_precondition_errors_raise_runtimeerror(cuda):
    from backtoreality_b200 import _ext
    x = torch.rand(1, 64, 3, device=cuda)
    with pytest.raises(RuntimeError, match="contiguous"):
        _ext.furthest_point_sampling(x.transpose(1, 2), 4)
    with pytest.raises(RuntimeError, match="float tensor"):
        _ext.furthest_point_sampling(x.double(), 4)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        _ext.furthest_point_sampling(x.cpu(), 4)
    f = torch.rand(1, 3, 64, device=cuda)
    with pytest.raises(RuntimeError, match="int tensor"):
        _ext.gather_points(f, torch.zeros(1, 4, dtype=torch.int64, device=cuda))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        _ext.gather_points(f, torch.zeros(1, 4, dtype=torch.int32))
    with pytest.raises(RuntimeError):
        _ext.furthest_point_sampling(torch.rand(1, 300000, 3, device=cuda), 4)  # beyond capacity


# ------------------------------------------------------------------ nn_distance -------------
@pytest.mark.parametrize("B,N,M", [(8, 256, 64), (8, 1024, 64), (4096, 1, 3), (2, 700, 1300)])
@pytest.mark.parametrize("kw", [{}, {"l1": True}, {"l1smooth": True, "delta": 0.3}])
def test_nn_distance_matches_reference_formulation(cuda, B, N, M, kw):
    """nn_distance (reference utils/nn_distance.py:34-61, the call shapes of
    loss_helper.py:65,100,131,180) -- values, indices and gradients against the reference's own
    tile-and-min formulation evaluated by torch in fp64/fp32 on the same inputs."""
    from backtoreality_b200 import nn_distance as nd
    g = torch.Generator().manual_seed(B + N + M)
    p1 = torch.rand(B, N, 3, generator=g).to(cuda).requires_grad_(True)
    p2 = torch.rand(B, M, 3, generator=g).to(cuda).requires_grad_(True)
    d1, i1, d2, i2 = nd.nn_distance(p1, p2, **kw)
    assert i1.dtype == torch.int64 and d1.shape == (B, N) and d2.shape == (B, M)
    (d1.sum() + 2.0 * d2.sum()).backward()
    g1, g2 = p1.grad.clone(), p2.grad.clone()
    q1 = p1.detach().clone().requires_grad_(True)
    q2 = p2.detach().clone().requires_grad_(True)
    diff = q1.unsqueeze(2).repeat(1, 1, M, 1) - q2.unsqueeze(1).repeat(1, N, 1, 1)
    cost = nd._pair_cost(diff, kw.get("l1smooth", False), kw.get("delta", 1.0), kw.get("l1", False))
    w1, j1 = torch.min(cost, dim=2)
    w2, j2 = torch.min(cost, dim=1)
    (w1.sum() + 2.0 * w2.sum()).backward()
    assert torch.allclose(d1, w1, rtol=1e-6, atol=1e-7) and torch.allclose(d2, w2, rtol=1e-6, atol=1e-7)
    # indices: identical except where two candidates tie to the last ulp
    assert float((i1 != j1).float().mean()) < 1e-3 and float((i2 != j2).float().mean()) < 1e-3
    assert rel_l2(g1.cpu().numpy(), q1.grad.cpu().numpy()) < 1e-4
    assert rel_l2(g2.cpu().numpy(), q2.grad.cpu().numpy()) < 1e-4


def test_fps_presorted_equals_one_call(cuda):
    """b2r_fps_sort + b2r_fps_ws_presorted (the sort run ahead of time by the pipelined step) give
    the indices of the single call; single-CTA scenes ignore the workspace."""
    from backtoreality_b200 import _ext
    for N, npnt in ((40000, 700), (3000, 200)):
        xyz = _t(scenes.batch(33, 2, N, C=0, kind="room", dup=0.2)[..., :3], cuda)
        ws = _ext.fps_presort(xyz)
        a = _ext.furthest_point_sampling(xyz, npnt, cluster=4, presorted=ws)
        b = _ext.furthest_point_sampling(xyz, npnt)
        assert torch.equal(a, b)
