"""CPU checks of the oracle itself (no GPU).

The C oracle restates the reference kernels literally (lane-by-lane FPS, serial ball query ...).
Here it is checked against INDEPENDENT numpy formulations on inputs quantised to a coarse binary
grid, where every product/sum is exact in fp32 (so contraction order cannot matter) and exact
ties are everywhere -- the cases the tie rules exist for.  The closed-form FPS tie rule used by the
CUDA kernel (max distance, then minimal (bitrev(k mod bs), k div bs)) is validated here against
the literal emulation.
"""
import numpy as np
import pytest

from oracle import cpu_ops


def _grid_points(rng, shape, span=4.0, step=1 / 64.0):
    return (rng.integers(0, int(span / step), size=shape).astype(np.float32) * np.float32(step))


def _bitrev(v, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (v & 1)
        v >>= 1
    return r


def _fps_closed_form(p, npoint):
    """FPS with the closed-form tie rule (what fps.cu implements)."""
    n = p.shape[0]
    bs = cpu_ops.block_threads(n)
    L = bs.bit_length() - 1
    k = np.arange(n)
    rank = np.array([(_bitrev(int(i) % bs, L) << 22) | (int(i) // bs) for i in k], dtype=np.int64)
    mag = (p.astype(np.float64) ** 2).sum(1)  # exact on the grid
    valid = ~(mag <= 1e-3)
    tmp = np.full(n, 1e10, np.float32)
    out = np.zeros(npoint, np.int32)
    old = 0
    for j in range(1, npoint):
        d = ((p - p[old]) ** 2).sum(1).astype(np.float32)
        tmp = np.where(valid, np.minimum(d, tmp), tmp)
        if not valid.any():
            old = 0
        else:
            cand = np.where(valid, tmp, -1.0)
            best = cand.max()
            ties = np.flatnonzero(cand == best)
            old = int(ties[np.argmin(rank[ties])])
        out[j] = old
    return out


@pytest.mark.parametrize("n,npoint", [(1, 1), (2, 2), (7, 7), (33, 20), (200, 64), (512, 100),
                                      (700, 128), (1500, 200), (5000, 64)])
def test_fps_literal_emulation_equals_closed_form_rule(n, npoint):
    rng = np.random.default_rng(n)
    p = _grid_points(rng, (n, 3), span=2.0, step=1 / 8.0)  # very coarse: massive ties
    p[::5] = 0.0                                             # points inside the |p|^2<=1e-3 hole
    got = cpu_ops.fps(p[None], npoint)[0]
    want = _fps_closed_form(p, npoint)
    assert np.array_equal(got, want)


def test_fps_all_invalid_and_first_index():
    p = np.zeros((1, 100, 3), np.float32)
    assert np.array_equal(cpu_ops.fps(p, 10), np.zeros((1, 10), np.int32))
    rng = np.random.default_rng(0)
    p = rng.random((3, 50, 3), dtype=np.float32) + 1
    assert (cpu_ops.fps(p, 5)[:, 0] == 0).all()


def test_block_threads_rule():
    # include/cuda_utils.h:20-24
    for n, want in [(1, 1), (2, 2), (3, 2), (255, 128), (256, 256), (511, 256), (512, 512),
                    (513, 512), (1024, 512), (2048, 512), (40000, 512)]:
        assert cpu_ops.block_threads(n) == want


@pytest.mark.parametrize("n,m,ns,r", [(300, 17, 8, 0.5), (1000, 33, 16, 0.25), (64, 5, 4, 1.0),
                                      (10, 3, 64, 100.0)])
def test_ball_query_vs_numpy(n, m, ns, r):
    rng = np.random.default_rng(n + m)
    xyz = _grid_points(rng, (2, n, 3))
    new = _grid_points(rng, (2, m, 3))
    got = cpu_ops.ball_query(new, xyz, r, ns)
    r2 = np.float32(r) * np.float32(r)
    for b in range(2):
        for j in range(m):
            d2 = ((new[b, j] - xyz[b]) ** 2).sum(1).astype(np.float32)
            hits = np.flatnonzero(d2 < r2)[:ns]
            want = np.zeros(ns, np.int32)
            if hits.size:
                want[:] = hits[0]
                want[:hits.size] = hits
            assert np.array_equal(got[b, j], want)


@pytest.mark.parametrize("n,m", [(50, 1), (50, 2), (50, 3), (200, 40), (64, 300)])
def test_three_nn_vs_numpy(n, m):
    rng = np.random.default_rng(n * m)
    unk = _grid_points(rng, (2, n, 3), step=1 / 4.0)
    kn = _grid_points(rng, (2, m, 3), step=1 / 4.0)
    d2, idx = cpu_ops.three_nn(unk, kn)
    for b in range(2):
        for j in range(n):
            d = ((unk[b, j] - kn[b]) ** 2).sum(1).astype(np.float32)
            order = np.lexsort((np.arange(m), d))[:3]     # by distance, then index
            wi = np.zeros(3, np.int32)
            wd = np.full(3, np.inf, np.float32)
            wi[:order.size] = order
            wd[:order.size] = d[order]
            assert np.array_equal(idx[b, j], wi)
            assert np.array_equal(d2[b, j], wd)


def test_movers_vs_numpy():
    rng = np.random.default_rng(5)
    B, C, N, NP, NS = 2, 6, 40, 9, 4
    f = rng.standard_normal((B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, (B, NP, NS)).astype(np.int32)
    out = cpu_ops.group(f, idx)
    for b in range(B):
        assert np.array_equal(out[b], f[b][:, idx[b]])
    g = rng.standard_normal((B, C, NP, NS)).astype(np.float32)
    want = np.zeros((B, C, N), np.float64)
    for b in range(B):
        for c in range(C):
            np.add.at(want[b, c], idx[b].ravel(), g[b, c].ravel())
    np.testing.assert_allclose(cpu_ops.group_grad(g, idx, N), want, rtol=1e-5, atol=1e-5)
    i1 = idx[:, :, 0].copy()
    assert np.array_equal(cpu_ops.gather(f, i1), np.stack([f[b][:, i1[b]] for b in range(B)]))
    g1 = rng.standard_normal((B, C, NP)).astype(np.float32)
    want = np.zeros((B, C, N), np.float64)
    for b in range(B):
        for c in range(C):
            np.add.at(want[b, c], i1[b], g1[b, c])
    np.testing.assert_allclose(cpu_ops.gather_grad(g1, i1, N), want, rtol=1e-5, atol=1e-5)


def test_three_interpolate_known_answer():
    """The reference's only fixture (pointnet2_test.py:18-30): idx [[0,1,2],[1,2,3]],
    weight [[1,1,1],[2,2,2]] on feats (1,2,4)."""
    feats = np.array([[[1.0, 2.0, 3.0, 4.0], [-1.0, 0.5, 0.25, 8.0]]], np.float32)
    idx = np.array([[[0, 1, 2], [1, 2, 3]]], np.int32)
    w = np.array([[[1, 1, 1], [2, 2, 2]]], np.float32)
    out = cpu_ops.interp(feats, idx, w)
    assert np.array_equal(out, np.array([[[6.0, 18.0], [-0.25, 17.5]]], np.float32))
    g = np.array([[[1.0, 10.0], [100.0, 1000.0]]], np.float32)
    gf = cpu_ops.interp_grad(g, idx, w, 4)
    assert np.array_equal(gf, np.array([[[1, 21, 21, 20], [100, 2100, 2100, 2000]]], np.float32))
