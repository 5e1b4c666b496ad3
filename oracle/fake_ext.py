"""A CPU stand-in for the reference's pybind module `pointnet2._ext`.

TEST INFRASTRUCTURE ONLY.  Exposes the nine functions registered in the reference's
`_ext_src/src/bindings.cpp:11-24` with the same names, argument order and return types, but on
CPU torch tensors, backed by the plain-C oracle (oracle/cpu_ops.py).  Used to

  * run the reference's UNMODIFIED Python stack (pointnet2_utils / pointnet2_modules /
    backbone_module) on CPU inside the build container, to generate golden fixtures
    (tests/golden/make_golden.py), and
  * drive this repo's own CPU baseline (oracle/cpu_backbone.py) from bench.py.

The reference's precondition checks (`_ext_src/include/utils.h:10-30`) are reproduced as
RuntimeError so misuse shows up the same way.
"""
import numpy as np
import torch

from . import cpu_ops


def _chk_f(t, name):
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be a float tensor")


def _chk_i(t, name):
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != torch.int32:
        raise RuntimeError(f"{name} must be an int tensor")


def _np(t):
    return t.detach().numpy()


def furthest_point_sampling(points, nsamples):
    _chk_f(points, "points")
    return torch.from_numpy(cpu_ops.fps(_np(points), int(nsamples)))


def gather_points(points, idx):
    _chk_f(points, "points")
    _chk_i(idx, "idx")
    return torch.from_numpy(cpu_ops.gather(_np(points), _np(idx)))


def gather_points_grad(grad_out, idx, n):
    _chk_f(grad_out, "grad_out")
    _chk_i(idx, "idx")
    return torch.from_numpy(cpu_ops.gather_grad(_np(grad_out), _np(idx), int(n)))


def ball_query(new_xyz, xyz, radius, nsample):
    _chk_f(new_xyz, "new_xyz")
    _chk_f(xyz, "xyz")
    return torch.from_numpy(
        cpu_ops.ball_query(_np(new_xyz), _np(xyz), float(np.float32(radius)), int(nsample)))


def group_points(points, idx):
    _chk_f(points, "points")
    _chk_i(idx, "idx")
    return torch.from_numpy(cpu_ops.group(_np(points), _np(idx)))


def group_points_grad(grad_out, idx, n):
    _chk_f(grad_out, "grad_out")
    _chk_i(idx, "idx")
    return torch.from_numpy(cpu_ops.group_grad(_np(grad_out), _np(idx), int(n)))


def three_nn(unknowns, knows):
    _chk_f(unknowns, "unknowns")
    _chk_f(knows, "knows")
    d2, idx = cpu_ops.three_nn(_np(unknowns), _np(knows))
    return [torch.from_numpy(d2), torch.from_numpy(idx)]


def three_interpolate(points, idx, weight):
    _chk_f(points, "points")
    _chk_i(idx, "idx")
    _chk_f(weight, "weight")
    return torch.from_numpy(cpu_ops.interp(_np(points), _np(idx), _np(weight)))


def three_interpolate_grad(grad_out, idx, weight, m):
    _chk_f(grad_out, "grad_out")
    _chk_i(idx, "idx")
    _chk_f(weight, "weight")
    return torch.from_numpy(cpu_ops.interp_grad(_np(grad_out), _np(idx), _np(weight), int(m)))
