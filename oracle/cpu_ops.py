"""ctypes front-end of oracle/liborc.so (the plain-C restatement in b2r_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of b2r_oracle.c.  numpy in, numpy out.
Function <-> reference kernel mapping is documented there (file:line).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Compile liborc.so with gcc (seconds).  Building the checker is not using it."""
    src = os.path.join(_HERE, "b2r_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_block_threads.restype = ctypes.c_int
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)


def block_threads(n):
    return int(lib().orc_block_threads(ctypes.c_int(int(n))))


def num_threads():
    return int(lib().orc_num_threads())


def fps(xyz, npoint):
    xyz, p = _f(xyz)
    B, N, _ = xyz.shape
    out = np.zeros((B, npoint), np.int32)
    lib().orc_fps(p, B, N, int(npoint), out.ctypes.data_as(_i32p))
    return out


def gather(features, idx):
    features, pf = _f(features)
    idx, pi = _i(idx)
    B, C, N = features.shape
    M = idx.shape[1]
    out = np.zeros((B, C, M), np.float32)
    lib().orc_gather(pf, pi, B, C, N, M, out.ctypes.data_as(_f32p))
    return out


def gather_grad(grad_out, idx, N):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    B, C, M = grad_out.shape
    out = np.zeros((B, C, N), np.float32)
    lib().orc_gather_grad(pg, pi, B, C, int(N), M, out.ctypes.data_as(_f32p))
    return out


def group(features, idx):
    features, pf = _f(features)
    idx, pi = _i(idx)
    B, C, N = features.shape
    _, NP, NS = idx.shape
    out = np.zeros((B, C, NP, NS), np.float32)
    lib().orc_group(pf, pi, B, C, N, NP, NS, out.ctypes.data_as(_f32p))
    return out


def group_grad(grad_out, idx, N):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    B, C, NP, NS = grad_out.shape
    out = np.zeros((B, C, N), np.float32)
    lib().orc_group_grad(pg, pi, B, C, int(N), NP, NS, out.ctypes.data_as(_f32p))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    new_xyz, pq = _f(new_xyz)
    xyz, px = _f(xyz)
    B, M, _ = new_xyz.shape
    N = xyz.shape[1]
    out = np.zeros((B, M, nsample), np.int32)
    lib().orc_ball_query(pq, px, B, N, M, ctypes.c_float(radius), int(nsample),
                         out.ctypes.data_as(_i32p))
    return out


def three_nn(unknown, known):
    unknown, pu = _f(unknown)
    known, pk = _f(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = np.zeros((B, n, 3), np.float32)
    idx = np.zeros((B, n, 3), np.int32)
    lib().orc_three_nn(pu, pk, B, n, m, d2.ctypes.data_as(_f32p), idx.ctypes.data_as(_i32p))
    return d2, idx


def interp(features, idx, weight):
    features, pf = _f(features)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    B, C, m = features.shape
    n = idx.shape[1]
    out = np.zeros((B, C, n), np.float32)
    lib().orc_interp(pf, pi, pw, B, C, m, n, out.ctypes.data_as(_f32p))
    return out


def interp_grad(grad_out, idx, weight, m):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    B, C, n = grad_out.shape
    out = np.zeros((B, C, int(m)), np.float32)
    lib().orc_interp_grad(pg, pi, pw, B, C, n, int(m), out.ctypes.data_as(_f32p))
    return out
