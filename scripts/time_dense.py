"""CUDA-event timing of the dense tcgen05 layers (csrc/dense.cu) at the VoteNet head shapes (B = 8):
forward, input gradient and weight gradient, each launch alone after an L2 flush.
   python scripts/time_dense.py [iters]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backtoreality_b200 import _ext, _lib  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
lib = _lib.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SHAPES = [("fp1.l0", 4096, 512, 256), ("fp1.l1", 4096, 256, 256), ("fp2.l0", 8192, 512, 256),
          ("fp2.l1", 8192, 256, 256), ("vgen.c1", 8192, 256, 256), ("vgen.c3", 8192, 256, 259),
          ("pnet.c1", 2048, 128, 128), ("pnet.c3", 2048, 128, 117)]


def timed(fn):
    ts = []
    for _ in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts[2:])[len(ts[2:]) // 2]


tot = [0.0, 0.0, 0.0]
for name, M, Cin, Cout in SHAPES:
    x = torch.randn(M, Cin, device=dev)
    w = torch.randn(Cout, Cin, device=dev) / Cin ** 0.5
    ld = (Cout + 3) // 4 * 4
    sc = torch.rand((Cin + 3) // 4 * 4, device=dev) + 0.5
    sh = torch.randn((Cin + 3) // 4 * 4, device=dev) * 0.1
    wi = torch.empty(lib.b2r_dense_image_bytes(Cout, Cin) // 4, device=dev)
    wt = torch.empty(lib.b2r_dense_image_bytes(Cin, Cout) // 4, device=dev)
    _lib.check(lib.b2r_dense_pack(P(w), Cout, Cin, P(wi), P(wt), _ext._stream()), "pack")
    z = torch.empty(M, ld, device=dev)
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
    d = _lib.DenseLayer()
    d.M, d.Cin, d.Cout, d.in_, d.ld_in, d.sc_in, d.sh_in = M, Cin, Cout, P(x), Cin, P(sc), P(sh)
    d.w_img, d.z, d.ld_z, d.stats = P(wi), P(z), ld, P(stats)
    t_f = timed(lambda: _lib.check(lib.b2r_dense_fwd(ctypes.byref(d), _ext._stream()), "fwd"))
    g = torch.randn(M, ld, device=dev)
    ca, cb, cc = (torch.randn(ld, device=dev) for _ in range(3))
    gin = torch.empty(M, Cin, device=dev)
    st_in = torch.zeros(2 * Cin, dtype=torch.float64, device=dev)
    dW = torch.zeros(Cout, Cin, device=dev)
    b = _lib.DenseLayerBwd()
    b.M, b.Cin, b.Cout, b.in_, b.ld_in, b.sc_in, b.sh_in = M, Cin, Cout, P(x), Cin, P(sc), P(sh)
    b.g, b.zz, b.ld_g, b.ca, b.cb, b.cc = P(g), P(z), ld, P(ca), P(cb), P(cc)
    b.wt_img, b.gin, b.ld_gin, b.stats_in = P(wt), P(gin), Cin, P(st_in)
    t_d = timed(lambda: _lib.check(lib.b2r_dense_bwd(ctypes.byref(b), _ext._stream()), "dgrad"))
    b.gin, b.dW = None, P(dW)
    t_w = timed(lambda: _lib.check(lib.b2r_dense_bwd(ctypes.byref(b), _ext._stream()), "wgrad"))
    fl = 2.0 * M * Cin * Cout
    print("%-8s M=%5d %3d->%3d   fwd %6.1f us (%5.1f TF/s)   dgrad %6.1f us   wgrad %6.1f us"
          % (name, M, Cin, Cout, t_f, fl / t_f / 1e6, t_d, t_w))
    tot[0] += t_f; tot[1] += t_d; tot[2] += t_w
print("sum over the 8 shapes: fwd %.1f  dgrad %.1f  wgrad %.1f us" % tuple(tot))
