"""Pins the oracle -- and the product -- against the REFERENCE'S OWN CUDA KERNELS.

oracle/_ref/_ext.so is the reference extension compiled for sm_100a straight from
/root/reference/.../_ext_src by oracle/Makefile (`make ref`); it is built in the container and
shipped to the GPU box with the snapshot.  Here the reference runs on the GPU and must agree
bit-for-bit with the C restatement (FPS / ball query / three_nn indices, movers) -- this is what
makes the oracle "pinned" rather than merely self-consistent -- and with libb2r.
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from backtoreality_b200 import scenes
from oracle import cpu_ops

pytestmark = pytest.mark.gpu
_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                   "_ext.so")


@pytest.fixture(scope="module")
def ref(cuda):
    if not os.path.isfile(_SO):
        pytest.skip("oracle/_ref/_ext.so not built (make -C oracle ref)")
    spec = importlib.util.spec_from_file_location("_ext", _SO)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _scene_xyz(i, B, N, kind="room", dup=0.2):
    return scenes.batch(i, B, N, C=0, kind=kind, dup=dup)[..., :3].copy()


@pytest.mark.parametrize("N,npoint,kind,dup", [(20000, 2048, "room", 0.2), (40000, 2048, "room", 0.2),
                                               (50000, 2048, "room", 0.4), (2048, 1024, "room_shifted", 0.2),
                                               (300, 300, "room", 0.5), (1000, 600, "uniform", 0.0),
                                               (513, 200, "room", 0.9)])
def test_fps_reference_kernel_vs_oracle_vs_product(cuda, ref, N, npoint, kind, dup):
    from backtoreality_b200 import _ext
    xyz = _scene_xyz(1, 2, N, kind, dup)
    x = torch.from_numpy(xyz).to(cuda)
    r = ref.furthest_point_sampling(x, npoint).cpu().numpy()
    assert np.array_equal(r, cpu_ops.fps(xyz, npoint)), "oracle != reference kernel"
    assert np.array_equal(r, _ext.furthest_point_sampling(x, npoint).cpu().numpy()), \
        "libb2r != reference kernel"


def test_fps_hole_points_reference_vs_oracle(cuda, ref):
    rng = np.random.default_rng(2)
    xyz = (rng.random((2, 5000, 3), dtype=np.float32) - 0.5) * 3
    xyz[:, ::9] *= 0.005
    xyz[0, 0] = 0
    x = torch.from_numpy(xyz).to(cuda)
    r = ref.furthest_point_sampling(x, 256).cpu().numpy()
    assert np.array_equal(r, cpu_ops.fps(xyz, 256))


@pytest.mark.parametrize("N,M,r,ns", [(40000, 2048, 0.2, 64), (2048, 1024, 0.4, 32), (1024, 256, 0.3, 16)])
def test_ball_query_reference_vs_oracle_vs_product(cuda, ref, N, M, r, ns):
    from backtoreality_b200 import _ext
    xyz = _scene_xyz(4, 2, N)
    new = np.take_along_axis(xyz, cpu_ops.fps(xyz, M)[..., None].astype(np.int64), axis=1)
    x, q = torch.from_numpy(xyz).to(cuda), torch.from_numpy(new).to(cuda)
    rr = ref.ball_query(q, x, r, ns).cpu().numpy()
    assert np.array_equal(rr, cpu_ops.ball_query(new, xyz, float(np.float32(r)), ns))
    assert np.array_equal(rr, _ext.ball_query(q, x, r, ns).cpu().numpy())


def test_three_nn_and_interpolate_reference_vs_oracle_vs_product(cuda, ref):
    from backtoreality_b200 import _ext
    xyz = _scene_xyz(6, 2, 4096)
    unk, kn = xyz[:, :1024].copy(), xyz[:, 1024:1536].copy()
    kn[:, 7] = kn[:, 3]
    u, k = torch.from_numpy(unk).to(cuda), torch.from_numpy(kn).to(cuda)
    rd, ri = ref.three_nn(u, k)
    od, oi = cpu_ops.three_nn(unk, kn)
    assert np.array_equal(ri.cpu().numpy(), oi) and np.array_equal(rd.cpu().numpy(), od)
    pd, pi = _ext.three_nn(u, k)
    assert torch.equal(pi, ri) and torch.equal(pd, rd)
    f = torch.randn(2, 64, 512, device=cuda)
    w = torch.rand(2, 1024, 3, device=cuda)
    ro = ref.three_interpolate(f, ri, w)
    assert torch.equal(ro, _ext.three_interpolate(f, ri, w))
    assert np.array_equal(ro.cpu().numpy(), cpu_ops.interp(f.cpu().numpy(), oi, w.cpu().numpy()))


def test_known_answer_on_reference_kernel_and_product(cuda, ref):
    """pointnet2_test.py:18-30 inputs; expected values computed by hand."""
    from backtoreality_b200 import _ext
    feats = torch.tensor([[[1.0, 2.0, 3.0, 4.0], [-1.0, 0.5, 0.25, 8.0]]], device=cuda)
    idx = torch.tensor([[[0, 1, 2], [1, 2, 3]]], dtype=torch.int32, device=cuda)
    w = torch.tensor([[[1.0, 1, 1], [2, 2, 2]]], device=cuda)
    want = torch.tensor([[[6.0, 18.0], [-0.25, 17.5]]], device=cuda)
    assert torch.equal(ref.three_interpolate(feats, idx, w), want)
    assert torch.equal(_ext.three_interpolate(feats, idx, w), want)
