"""The wide point-major dense layers (csrc/dense.cu) and their glue (csrc/heads.cu) against fp64
torch references, then the modules built on them against the reference formulation (torch /
cuDNN in fp32): PointnetFPModule's SharedMLP, VotingModule + normalisation, ProposalModule's head
(reference pointnet2_modules.py:505-514, models/voting_module.py:38-65, votenet.py:93-94,
models/proposal_module.py:115-119).

Tolerance: TF32 operands (10-bit mantissa), FP32 accumulate, forward and backward: rel-L2 <= 2e-3
per GEMM against fp64.
"""
import ctypes

import numpy as np
import pytest
import torch

from _util import emu_log, rel_l2, round_tf32

pytestmark = pytest.mark.gpu
TOL = 2e-3
# against the same products with the operands rounded as the kernel rounds them (cvt.rna.tf32 of x,
# W, dz; fp64 products and sums): what is left is fp32 accumulation (measured <= 1.9e-6; profiles/r02/emulation_parity.log)
EMU_TOL = float(__import__("os").environ.get("B2R_EMU_TOL_DENSE", "1e-5"))


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _pad4(v):
    out = torch.zeros((v.numel() + 3) // 4 * 4, dtype=torch.float32, device=v.device)
    out[:v.numel()] = v
    return out


def _pack(w):
    from backtoreality_b200 import _ext, _lib
    lib = _lib.lib()
    Cout, Cin = w.shape
    wi = torch.empty(lib.b2r_dense_image_bytes(Cout, Cin) // 4, device=w.device)
    wt = torch.empty(lib.b2r_dense_image_bytes(Cin, Cout) // 4, device=w.device)
    _lib.check(lib.b2r_dense_pack(_p(w), Cout, Cin, _p(wi), _p(wt), _ext._stream()), "pack")
    return wi, wt


@pytest.mark.parametrize("M,Cin,Cout,pro,bias", [(4096, 512, 256, False, False), (8192, 256, 259, True, True),
                                                (2048, 128, 117, True, True), (1000, 132, 24, True, False),
                                                (130, 4, 300, False, True), (8192, 256, 256, True, True)])
def test_dense_forward(cuda, M, Cin, Cout, pro, bias):
    from backtoreality_b200 import _ext, _lib
    g = torch.Generator(device="cpu").manual_seed(M + Cin + Cout)
    ld_in = Cin + 4
    xin = torch.randn(M, ld_in, generator=g).to(cuda)
    w = (torch.randn(Cout, Cin, generator=g) / Cin ** 0.5).to(cuda)
    sc = (torch.rand(Cin, generator=g) + 0.5).to(cuda)
    sc[::3] *= -1
    sh = (torch.randn(Cin, generator=g) * 0.3).to(cuda)
    b = torch.randn(Cout, generator=g).to(cuda) if bias else None
    wi, _ = _pack(w)
    ld_z = (Cout + 3) // 4 * 4
    z = torch.full((M, ld_z), float("nan"), device=cuda)
    stats = torch.zeros(2, Cout, dtype=torch.float64, device=cuda)
    d = _lib.DenseLayer()
    d.M, d.Cin, d.Cout = M, Cin, Cout
    d.in_, d.ld_in = _p(xin), ld_in
    scp, shp = _pad4(sc), _pad4(sh)
    if pro:
        d.sc_in, d.sh_in = _p(scp), _p(shp)
    d.w_img, d.bias, d.z, d.ld_z, d.stats = _p(wi), _p(b), _p(z), ld_z, _p(stats)
    _lib.check(_lib.lib().b2r_dense_fwd(ctypes.byref(d), _ext._stream()), "dense_fwd")
    torch.cuda.synchronize()
    x = xin[:, :Cin].double()
    if pro:
        x = torch.relu(x * sc.double() + sh.double())
    want = x @ w.double().t()
    if bias:
        want = want + b.double()
    got = z[:, :Cout]
    assert bool(torch.isfinite(got).all())
    assert rel_l2(got.cpu().numpy(), want.cpu().numpy()) < TOL
    want_e = round_tf32(x.float()).double() @ round_tf32(w).double().t()
    if bias:
        want_e = want_e + b.double()
    e = rel_l2(got.cpu().numpy(), want_e.cpu().numpy())
    emu_log("heads_dense_fwd M=%d %dx%d" % (M, Cin, Cout), z=e)
    assert e < EMU_TOL, e
    np.testing.assert_allclose(stats[0].cpu().numpy(), got.double().sum(0).cpu().numpy(), rtol=1e-6, atol=1e-3)
    np.testing.assert_allclose(stats[1].cpu().numpy(), (got.double() ** 2).sum(0).cpu().numpy(), rtol=1e-6, atol=1e-3)


@pytest.mark.parametrize("M,Cin,Cout,form,masked", [(4096, 512, 256, "bn", False), (8192, 256, 259, "direct", True),
                                                   (2048, 128, 117, "direct", True), (1000, 132, 24, "bn", True),
                                                   (8192, 256, 256, "bn", True), (130, 8, 300, "bn", False),
                                                   (2048, 128, 128, "bn", True)])
def test_dense_backward(cuda, M, Cin, Cout, form, masked):
    """gin = (dz W) * mask with the BatchNorm-backward sums, and dW = dz^T x, against fp64."""
    from backtoreality_b200 import _ext, _lib
    g = torch.Generator(device="cpu").manual_seed(M * 3 + Cin + Cout)
    ld_in, ld_g = Cin, (Cout + 3) // 4 * 4
    xin = torch.randn(M, ld_in, generator=g).to(cuda)
    w = (torch.randn(Cout, Cin, generator=g) / Cin ** 0.5).to(cuda)
    sc = (torch.rand(Cin, generator=g) + 0.5).to(cuda)
    sc[::5] *= -1
    sh = (torch.randn(Cin, generator=g) * 0.3).to(cuda)
    gr = torch.randn(M, ld_g, generator=g).to(cuda)
    zz = torch.randn(M, ld_g, generator=g).to(cuda)
    ca = (torch.rand(Cout, generator=g) + 0.5).to(cuda)
    cb = (torch.randn(Cout, generator=g) * 0.2).to(cuda)
    cc = (torch.randn(Cout, generator=g) * 0.1).to(cuda)
    _, wt = _pack(w)
    gin = torch.full((M, Cin), float("nan"), device=cuda)
    stats = torch.zeros(2, Cin, dtype=torch.float64, device=cuda)
    dW = torch.ones(Cout, Cin, device=cuda)          # accumulated on top of what is there
    b = _lib.DenseLayerBwd()
    b.M, b.Cin, b.Cout = M, Cin, Cout
    b.in_, b.ld_in = _p(xin), ld_in
    scp, shp, cap, cbp, ccp = (_pad4(v) for v in (sc, sh, ca, cb, cc))
    if masked:
        b.sc_in, b.sh_in = _p(scp), _p(shp)
    b.g, b.zz, b.ld_g = _p(gr), _p(zz), ld_g
    if form == "bn":
        b.ca, b.cb, b.cc = _p(cap), _p(cbp), _p(ccp)
    b.wt_img, b.gin, b.ld_gin, b.stats_in, b.dW = _p(wt), _p(gin), Cin, _p(stats), _p(dW)
    _lib.check(_lib.lib().b2r_dense_bwd(ctypes.byref(b), _ext._stream()), "dense_bwd")
    torch.cuda.synchronize()
    dz = gr[:, :Cout].double()
    if form == "bn":
        dz = ca.double() * dz + cb.double() * zz[:, :Cout].double() + cc.double()
    x = xin.double()
    if masked:
        pre = torch.addcmul(sh, xin, sc)            # as the kernel evaluates the mask: one fp32 fma
        x = torch.relu(x * sc.double() + sh.double())
    want_dW = dz.t() @ x
    want_g = dz @ w.double()
    if masked:
        want_g = want_g * (pre > 0)
    assert rel_l2((dW - 1.0).cpu().numpy(), want_dW.cpu().numpy()) < TOL
    assert rel_l2(gin.cpu().numpy(), want_g.cpu().numpy()) < TOL
    dz_e = round_tf32(dz.float()).double()
    want_g_e = dz_e @ round_tf32(w).double()
    if masked:
        want_g_e = want_g_e * (pre > 0)
    e_w = rel_l2((dW - 1.0).cpu().numpy(), (dz_e.t() @ round_tf32(x.float()).double()).cpu().numpy())
    e_g = rel_l2(gin.cpu().numpy(), want_g_e.cpu().numpy())
    emu_log("heads_dense_bwd %s M=%d %dx%d" % (form, M, Cin, Cout), dW=e_w, g_in=e_g)
    assert e_w < EMU_TOL and e_g < EMU_TOL, (e_w, e_g)
    if masked:
        np.testing.assert_allclose(stats[0].cpu().numpy(), gin.double().sum(0).cpu().numpy(), rtol=1e-5, atol=1e-2)
        np.testing.assert_allclose(stats[1].cpu().numpy(), (gin.double() * xin.double()).sum(0).cpu().numpy(),
                                   rtol=1e-5, atol=1e-2)


def test_interp_cat_matches_three_interpolate_and_cat(cuda):
    from backtoreality_b200 import dense_mlp, pointnet2_utils
    from backtoreality_b200.pointnet2_modules import PointnetFPModule
    g = torch.Generator(device="cpu").manual_seed(5)
    B, n, m, C2, C1 = 2, 300, 77, 64, 32
    unknown = torch.rand(B, n, 3, generator=g).to(cuda)
    known = torch.rand(B, m, 3, generator=g).to(cuda)
    kf = torch.randn(B, C2, m, generator=g).to(cuda).requires_grad_(True)
    sf = torch.randn(B, C1, n, generator=g).to(cuda).requires_grad_(True)
    idx, weight = PointnetFPModule.interpolation_weights(unknown, known)
    want = torch.cat([pointnet2_utils.three_interpolate(kf, idx, weight), sf], dim=1)   # (B,C,n)
    patt = torch.sin(torch.arange(want.numel(), device=cuda, dtype=torch.float32)).view_as(want)
    (want * patt).sum().backward()
    gk0, gs0 = kf.grad.clone(), sf.grad.clone()
    kf_pm = kf.detach().transpose(1, 2).contiguous().requires_grad_(True)
    sf_pm = sf.detach().transpose(1, 2).contiguous().requires_grad_(True)
    got = dense_mlp.interp_cat(kf_pm, sf_pm, idx, weight)                               # (B*n, C)
    assert torch.equal(got.view(B, n, C2 + C1).transpose(1, 2), want.detach())          # same contraction
    (got.view(B, n, C2 + C1).transpose(1, 2) * patt).sum().backward()
    assert rel_l2(kf_pm.grad.transpose(1, 2).cpu().numpy(), gk0.cpu().numpy()) < 1e-5
    assert torch.equal(sf_pm.grad.transpose(1, 2), gs0)


def test_vote_tail_matches_torch(cuda):
    from backtoreality_b200 import dense_mlp
    g = torch.Generator(device="cpu").manual_seed(9)
    B, n, C = 2, 333, 256
    ld = 260
    net0 = torch.randn(B * n, ld, generator=g).to(cuda)
    sx0 = torch.rand(B, n, 3, generator=g).to(cuda)
    sf0 = torch.randn(B, n, C, generator=g).to(cuda)
    outs = []
    for fused in (True, False):
        net = net0.clone().requires_grad_(True)
        sx = sx0.clone().requires_grad_(True)
        sf = sf0.clone().requires_grad_(True)
        if fused:
            vx, vf = dense_mlp.vote_tail(net, sx, sf)
        else:
            nn_ = net.view(B, n, ld)
            vx = sx + nn_[..., 0:3]
            v = sf + nn_[..., 3:3 + C]
            vf = v / torch.norm(v, p=2, dim=2, keepdim=True)
        patt = torch.sin(torch.arange(vf.numel(), device=cuda, dtype=torch.float32)).view_as(vf)
        ((vf * patt).sum() + (vx * 0.37).sum()).backward()
        outs.append((vx.detach(), vf.detach(), net.grad, sx.grad, sf.grad))
    a, b = outs
    assert rel_l2(a[0].cpu().numpy(), b[0].cpu().numpy()) < 1e-6
    assert rel_l2(a[1].cpu().numpy(), b[1].cpu().numpy()) < 1e-6
    assert rel_l2(a[2][:, :3 + C].cpu().numpy(), b[2][:, :3 + C].cpu().numpy()) < 1e-5
    assert float(a[2][:, 3 + C:].abs().max()) == 0.0
    assert rel_l2(a[3].cpu().numpy(), b[3].cpu().numpy()) < 1e-6
    assert rel_l2(a[4].cpu().numpy(), b[4].cpu().numpy()) < 1e-5


def _fp32_convs():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    return old


def _randomize_bn(mod):
    for m in mod.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.weight.data = torch.randn_like(m.weight) * 0.5 + 0.8
            m.bias.data = torch.randn_like(m.bias) * 0.2
            m.running_mean.data = torch.randn_like(m.running_mean) * 0.1
            m.running_var.data = torch.rand_like(m.running_var) + 0.5
            m.momentum = 0.3


@pytest.mark.parametrize("training", [True, False])
def test_fp_module_dense_vs_torch(cuda, training, monkeypatch):
    """PointnetFPModule through csrc/dense.cu against the reference formulation (three_interpolate,
    cat, cuDNN fp32 SharedMLP): outputs, input gradients, parameter gradients, running stats."""
    import copy
    from backtoreality_b200 import dense_mlp
    from backtoreality_b200.pointnet2_modules import PointnetFPModule
    old = _fp32_convs()
    try:
        torch.manual_seed(4)
        fp = PointnetFPModule(mlp=[256 + 128, 256, 288]).to(cuda)
        _randomize_bn(fp)
        fp.train(training)
        ref = copy.deepcopy(fp)
        g = torch.Generator(device="cpu").manual_seed(2)
        B, n, m = 2, 700, 256
        unknown = torch.rand(B, n, 3, generator=g).to(cuda)
        known = torch.rand(B, m, 3, generator=g).to(cuda)
        uf0 = torch.randn(B, 128, n, generator=g).to(cuda)
        kf0 = torch.randn(B, 256, m, generator=g).to(cuda)
        outs = []
        for mod, dense in ((fp, True), (ref, False)):
            monkeypatch.setattr(dense_mlp, "ENABLED", dense)
            uf = uf0.clone().requires_grad_(True)
            kf = kf0.clone().requires_grad_(True)
            y = mod(unknown, known, uf, kf)
            patt = torch.sin(torch.arange(y.numel(), device=cuda, dtype=torch.float32) * 1.7).view_as(y)
            (y * patt).sum().backward()
            outs.append((y.detach(), uf.grad, kf.grad, [p.grad for p in mod.parameters()],
                         {k: v.clone() for k, v in mod.state_dict().items() if "running" in k or "tracked" in k}))
        (y1, gu1, gk1, gp1, rs1), (y0, gu0, gk0, gp0, rs0) = outs
        assert y1.shape == (B, 288, n)
        assert rel_l2(y1.cpu().numpy(), y0.cpu().numpy()) < 3e-3
        # TF32 through BN + ReLU twice: a few flipped ReLU masks (see test_mlp_gpu's block test)
        tol = 4e-2
        assert rel_l2(gu1.cpu().numpy(), gu0.cpu().numpy()) < tol
        assert rel_l2(gk1.cpu().numpy(), gk0.cpu().numpy()) < tol
        for (name, _), a, b in zip(fp.named_parameters(), gp1, gp0):
            assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < tol, name
        for k in rs0:
            assert rel_l2(rs1[k].float().cpu().numpy(), rs0[k].float().cpu().numpy()) < 3e-3, k
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("training", [True, False])
def test_vote_heads_dense_vs_torch(cuda, training, monkeypatch):
    """VotingModule (+ normalisation) and ProposalModule on the dense path against the reference's
    torch formulation with identical parameters: end_points, gradients of the seed inputs and of
    every vgen / pnet parameter (conv biases in front of BatchNorm included)."""
    import copy
    from backtoreality_b200 import dense_mlp
    from backtoreality_b200.votenet import ProposalModule, VotingModule
    old = _fp32_convs()
    try:
        torch.manual_seed(6)
        vgen = VotingModule(1, 256).to(cuda)
        # seed_fps: the proposals are sampled on the seed coordinates, which are identical in both
        # arms (FPS on the votes would amplify their 1e-3 TF32 differences into other samples)
        pnet = ProposalModule(22, 1, 22, np.ones((22, 3), np.float32), 64, "seed_fps").to(cuda)
        _randomize_bn(vgen)
        _randomize_bn(pnet)
        vgen.train(training)
        pnet.train(training)
        ref_v, ref_p = copy.deepcopy(vgen), copy.deepcopy(pnet)
        g = torch.Generator(device="cpu").manual_seed(8)
        B, n = 2, 512
        sx0 = (torch.rand(B, n, 3, generator=g) * 3.0).to(cuda)
        sf0 = torch.randn(B, 256, n, generator=g).to(cuda)
        outs = []
        for (v, p), dense in (((vgen, pnet), True), ((ref_v, ref_p), False)):
            monkeypatch.setattr(dense_mlp, "ENABLED", dense)
            sx = sx0.clone().requires_grad_(True)
            sf = sf0.clone().requires_grad_(True)
            res = v.forward_normalized(sx, sf) if dense else None
            if dense:
                assert res is not None
                vx, vf, vf_pm = res
            else:
                vx, vf = v(sx, sf)
                vf = vf.div(torch.norm(vf, p=2, dim=1).unsqueeze(1))
                vf_pm = None
            ep = p(vx, vf, {"seed_xyz": sx}, features_pm=vf_pm)
            raw = ep["proposal_scores_raw"]
            patt = torch.sin(torch.arange(raw.numel(), device=cuda, dtype=torch.float32) * 0.9).view(raw.shape)
            ((raw * patt).sum() + (vx * 0.21).sum() + ep["center"].sum() * 0.1).backward()
            outs.append((vx.detach(), vf.detach(), raw.detach().contiguous(), sx.grad, sf.grad,
                         [q.grad for q in list(v.parameters()) + list(p.parameters())]))
        a, b = outs
        assert rel_l2(a[0].cpu().numpy(), b[0].cpu().numpy()) < 3e-3
        assert rel_l2(a[1].cpu().numpy(), b[1].cpu().numpy()) < 3e-3
        assert a[2].shape == b[2].shape == (B, 2 + 3 + 2 + 22 * 4 + 22, 64)
        assert rel_l2(a[2].cpu().numpy(), b[2].cpu().numpy()) < 2e-2
        tol = 8e-2    # through five TF32 blocks (vgen, SA vote aggregation with BF16 backward, pnet)
        # the seed-xyz gradient also takes the ball-query / max-pool routing of the vote
        # aggregation, where a 1e-3 difference in a vote moves whole samples: measured 0.095
        assert rel_l2(a[3].cpu().numpy(), b[3].cpu().numpy()) < 2 * tol
        assert rel_l2(a[4].cpu().numpy(), b[4].cpu().numpy()) < 2 * tol
        names = [k for k, _ in vgen.named_parameters()] + [k for k, _ in pnet.named_parameters()]
        for name, ga, gb in zip(names, a[5], b[5]):
            assert (ga is None) == (gb is None), name
            if ga is None:
                continue
            if training and name in ("conv1.bias", "conv2.bias"):
                # a bias in front of a training-mode BatchNorm has a zero gradient analytically:
                # both arms only hold rounding noise there
                assert float(ga.abs().max()) < 1e-3 and float(gb.abs().max()) < 1e-3, name
                continue
            assert rel_l2(ga.cpu().numpy(), gb.cpu().numpy()) < 2 * tol, name
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
