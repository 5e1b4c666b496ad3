"""Module / backbone parity on the GPU: the product's drop-in API (pointnet2_utils autograd
functions, PointnetSAModuleVotes, PointnetFPModule, Pointnet2Backbone) against
  * the golden fixtures generated from the reference's own Python stack, and
  * the oracle's CPU port (oracle/cpu_modules.py) at BASELINE.json's full 40k-point size.

Tolerances: indices bit-exact; features rel-L2 <= 1e-4 with TF32 disabled (pure
fp32 MLP), <= 5e-3 in the default TF32 MLP mode the reference itself runs in.
Gradients through the whole backbone: rel-L2 <= 1e-2 (fp32) / 5e-2 (TF32).  That is not slack
for the kernels -- it is the reference's own noise floor: the fp32 CPU reference stack run with
1 and with 8 threads (different summation order only) disagrees with itself by 1e-3 .. 3e-3 on
every SA-layer weight gradient of these fixtures (train-mode BatchNorm backward amplifies forward
rounding differences of ~1e-6; measured with oracle/cpu_modules.py, see DESIGN.md "Tolerances").
Per-block gradient parity at tight tolerance is in test_fused_sa_gpu.py against fp64.
"""
import numpy as np
import pytest
import torch

from _util import emu_log, golden, pattern_like, rel_l2, sub, weight_checksum
from backtoreality_b200 import scenes

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
TF32_TOL = 5e-3
GRAD_TOL = 1e-2        # fp32 mode, see the module docstring
GRAD_TOL_TF32 = 8e-2   # one TF32 block vs fp32: see test_mlp_gpu.py / profiles/r01/tf32_gradient_noise*


@pytest.fixture(params=["unfused_fp32", "fused_tf32"])
def fp32_mlp(request):
    """Two arms.  unfused_fp32: QueryAndGroup kernel + cuDNN SharedMLP in true fp32 (tight
    tolerances; checks the movers and the module plumbing).  fused_tf32: the product's default
    path, the tcgen05 SA block with TF32 operands (the precision the reference's cuDNN runs at).
    Yields (feature tolerance, gradient tolerance) for a SINGLE block."""
    from backtoreality_b200 import fused_sa
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32,
           fused_sa.ENABLED)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    fused_sa.ENABLED = request.param == "fused_tf32"
    yield (FP32_TOL, GRAD_TOL) if request.param == "unfused_fp32" else (TF32_TOL, GRAD_TOL_TF32)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, fused_sa.ENABLED = old


def _set_arm(arm):
    """fp32: unfused + cuDNN fp32;  cudnn_tf32: unfused + cuDNN TF32 (what the reference runs);
    fused: the tcgen05 TF32 SA block."""
    from backtoreality_b200 import fused_sa
    tf32 = arm == "cudnn_tf32"
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
    fused_sa.ENABLED = arm == "fused"


def _backbone_errors(g, cuda, arm):
    from backtoreality_b200.backbone_module import Pointnet2Backbone
    _set_arm(arm)
    torch.manual_seed(int(g["seed"]))
    net = Pointnet2Backbone(input_feature_dim=int(g["C"]), fp2_out=int(g["fp2_out"]))
    assert abs(weight_checksum(net) - float(g["wsum"])) < 1e-6 * float(g["wsum"])
    net = net.to(cuda).train(bool(g["train"]))
    pc = torch.from_numpy(scenes.batch(50, int(g["B"]), int(g["N"]), C=int(g["C"]), kind="room",
                                       dup=0.2)).to(cuda)
    ep = net(pc)
    assert np.array_equal(ep["sa1_inds"].cpu().numpy(), g["sa1_inds"])
    assert np.array_equal(ep["sa2_inds"].cpu().numpy(), g["sa2_inds"])
    feat = {k: rel_l2(sub(ep[k]), g[k]) for k in ("sa1_features", "sa2_features", "sa3_features",
                                                   "sa4_features", "fp2_features")}
    (ep["fp2_features"] * pattern_like(ep["fp2_features"])).sum().backward()
    grad = {"g_sa1_l0": rel_l2(sub(net.sa1.mlp_module.layer0.conv.weight.grad), g["g_sa1_l0"]),
            "g_sa2_l0": rel_l2(sub(net.sa2.mlp_module.layer0.conv.weight.grad), g["g_sa2_l0"]),
            "g_sa4_l2": rel_l2(sub(net.sa4.mlp_module.layer2.conv.weight.grad), g["g_sa4_l2"]),
            "g_fp1_l0": rel_l2(sub(net.fp1.mlp.layer0.conv.weight.grad), g["g_fp1_l0"]),
            "g_fp2_l1_bn": rel_l2(sub(net.fp2.mlp.layer1.bn.bn.weight.grad), g["g_fp2_l1_bn"])}
    stats = {}
    if g["train"]:
        bn = net.sa1.mlp_module.layer0.bn.bn
        stats = {"rm": rel_l2(bn.running_mean.cpu().numpy(), g["rm_sa1_l0"]),
                 "rv": rel_l2(bn.running_var.cpu().numpy(), g["rv_sa1_l0"])}
    return feat, grad, stats


@pytest.mark.parametrize("fixture", ["backbone_votenet_eval.npz", "backbone_votenet_train.npz",
                                     "backbone_gf3d_train.npz"])
def test_backbone_vs_reference_python_golden(cuda, fixture):
    """Whole backbone, fwd + bwd, against fixtures generated from the reference's own Python.

    fp32 arm (unfused movers + cuDNN fp32): features 1e-4, gradients 1e-2 (module docstring).
    Product arm (fused tcgen05 TF32 SA blocks): TF32 rounding of ~5e-4 in the forward flips that
    fraction of ReLU masks and max-pool winners, and through four stacked SA blocks that is a
    10-25 % gradient L2 difference from fp32 -- for ANY TF32 implementation, including the cuDNN
    TF32 path the reference itself runs by default (profiles/r01/tf32_gradient_noise_cudnn_vs_
    fused.log).  So the product is held to "no further from the fp32 reference than cuDNN TF32
    is": err_fused <= 1.5 * err_cudnn_tf32 + floor, with both measured here on the same inputs.
    """
    from backtoreality_b200 import fused_sa
    g = golden(fixture)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        feat, grad, stats = _backbone_errors(g, cuda, "fp32")
        for k, e in feat.items():
            assert e < FP32_TOL, (k, e)
        for k, e in grad.items():
            assert e < GRAD_TOL, (k, e)
        for k, e in stats.items():
            assert e < FP32_TOL, (k, e)
        feat_c, grad_c, _ = _backbone_errors(g, cuda, "cudnn_tf32")
        feat_f, grad_f, stats_f = _backbone_errors(g, cuda, "fused")
        for k in feat_f:
            assert feat_f[k] < 1.5 * feat_c[k] + 1e-3, (k, feat_f[k], feat_c[k])
        for k in grad_f:
            assert grad_f[k] < 1.5 * grad_c[k] + 2e-2, (k, grad_f[k], grad_c[k])
        for k, e in stats_f.items():
            assert e < TF32_TOL, (k, e)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        fused_sa.ENABLED = True


def test_vote_aggregation_golden_xyz_gradients_and_given_inds(cuda, fp32_mlp):
    from backtoreality_b200.pointnet2_modules import PointnetSAModuleVotes
    g = golden("vote_aggregation.npz")
    ftol, gtol = fp32_mlp
    torch.manual_seed(int(g["seed"]))
    mlp = [32, 32, 32, 32]
    sa = PointnetSAModuleVotes(npoint=64, radius=0.3, nsample=16, mlp=mlp, use_xyz=True,
                               normalize_xyz=True)
    assert mlp[0] == 35  # the reference mutates the caller's list (pointnet2_modules.py:204-206)
    assert abs(weight_checksum(sa) - float(g["wsum"])) < 1e-6 * float(g["wsum"])
    sa = sa.to(cuda)
    gen = torch.Generator().manual_seed(int(g["seed"]) + 1)
    xyz = (torch.rand(2, 256, 3, generator=gen) * 2.0 + 0.5).to(cuda).requires_grad_(True)
    feats = torch.randn(2, 32, 256, generator=gen).to(cuda).requires_grad_(True)
    new_xyz, new_feats, inds = sa(xyz, feats)
    assert np.array_equal(inds.cpu().numpy(), g["inds"])
    assert np.array_equal(new_xyz.detach().cpu().numpy(), g["new_xyz"])
    assert rel_l2(new_feats.detach().cpu().numpy(), g["new_feats"]) < ftol
    ((new_feats * pattern_like(new_feats)).sum() + (new_xyz * 0.37).sum()).backward()
    assert rel_l2(xyz.grad.cpu().numpy(), g["g_xyz"]) < gtol
    assert rel_l2(sub(feats.grad), g["g_feats"]) < gtol
    # explicit inds + features=None (GroupFree3D SA1 style, proposal_module.py:97-100)
    torch.manual_seed(int(g["seed"]))
    sa2 = PointnetSAModuleVotes(npoint=64, radius=0.4, nsample=8, mlp=[0, 16, 16], use_xyz=True,
                                normalize_xyz=True)
    assert abs(weight_checksum(sa2) - float(g["wsum2"])) < 1e-6 * float(g["wsum2"])
    sa2 = sa2.to(cuda)
    given = torch.from_numpy(g["given"]).to(cuda)
    nx, nf, gi = sa2(xyz.detach(), None, given)
    assert torch.equal(gi, given)
    assert np.array_equal(nx.cpu().numpy(), g["nx2"])
    assert rel_l2(nf.detach().cpu().numpy(), g["nf2"]) < ftol


def test_fp_module_golden(cuda, fp32_mlp):
    from backtoreality_b200.pointnet2_modules import PointnetFPModule
    g = golden("fp_module.npz")
    ftol, gtol = fp32_mlp
    torch.manual_seed(int(g["seed"]))
    fp = PointnetFPModule(mlp=[48 + 16, 32, 24])
    assert abs(weight_checksum(fp) - float(g["wsum"])) < 1e-6 * float(g["wsum"])
    fp = fp.to(cuda)
    gen = torch.Generator().manual_seed(int(g["seed"]) + 1)
    unknown = torch.rand(2, 100, 3, generator=gen)
    known = torch.rand(2, 37, 3, generator=gen)
    known[:, 5] = known[:, 2]
    uf = torch.randn(2, 16, 100, generator=gen).to(cuda).requires_grad_(True)
    kf = torch.randn(2, 48, 37, generator=gen).to(cuda).requires_grad_(True)
    y = fp(unknown.to(cuda), known.to(cuda), uf, kf)
    assert rel_l2(y.detach().cpu().numpy(), g["y"]) < ftol
    (y * pattern_like(y)).sum().backward()
    assert rel_l2(kf.grad.cpu().numpy(), g["g_kf"]) < gtol
    assert rel_l2(sub(uf.grad), g["g_uf"]) < gtol


def test_reference_gradcheck_of_three_interpolate(cuda):
    """The reference's only test (pointnet2_test.py:18-30), against the product op."""
    from backtoreality_b200 import pointnet2_utils
    feats = torch.randn(1, 2, 4, requires_grad=True).float().cuda()

    def interpolate_func(inputs):
        idx = torch.from_numpy(np.array([[[0, 1, 2], [1, 2, 3]]])).int().cuda()
        weight = torch.from_numpy(np.array([[[1, 1, 1], [2, 2, 2]]])).float().cuda()
        return pointnet2_utils.three_interpolate(inputs, idx, weight)

    assert torch.autograd.gradcheck(interpolate_func, feats, atol=1e-1, rtol=1e-1)


@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_backbone_full_size_40k_vs_oracle_port(cuda, mode):
    """BASELINE.json config: 40k-point ScanNet-shaped scenes, train-mode BN, fwd + bwd, against
    the oracle's CPU port.  fp32 arm: absolute tolerances.  tf32 arm (the product's fused path):
    no further from the oracle than the reference's own arithmetic (unfused + cuDNN TF32) is."""
    from backtoreality_b200.backbone_module import Pointnet2Backbone
    from backtoreality_b200 import fused_sa
    from oracle import cpu_modules
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        torch.manual_seed(3)
        port = cpu_modules.Backbone(input_feature_dim=1).train()
        state = {k: v.clone() for k, v in port.state_dict().items()}
        pc = torch.from_numpy(scenes.batch(20, 2, 40000, C=1, kind="room", dup=0.2))
        want = port(pc)
        (want["fp2_features"] * pattern_like(want["fp2_features"])).sum().backward()
        feats = ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features")

        def run(arm):
            _set_arm(arm)
            net = Pointnet2Backbone(input_feature_dim=1)
            net.load_state_dict(state)
            net = net.to(cuda).train()
            got = net(pc.to(cuda))
            for k in ("sa1_inds", "sa2_inds", "fp2_inds"):
                assert torch.equal(got[k].cpu(), want[k]), k
            for k in ("sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz"):
                assert torch.equal(got[k].cpu(), want[k]), k
            fe = {k: rel_l2(got[k].detach().cpu().numpy(), want[k].detach().numpy()) for k in feats}
            (got["fp2_features"] * pattern_like(got["fp2_features"])).sum().backward()
            ge = {}
            for (n1, p1), (n2, p2) in zip(port.named_parameters(), net.named_parameters()):
                assert n1 == n2
                ge[n1] = rel_l2(p2.grad.cpu().numpy(), p1.grad.numpy())
            return fe, ge

        if mode == "fp32":
            fe, ge = run("fp32")
            for k, e in fe.items():
                assert e < FP32_TOL, (k, e)
            for k, e in ge.items():
                assert e < GRAD_TOL, (k, e)
        else:
            fe_c, ge_c = run("cudnn_tf32")
            fe_f, ge_f = run("fused")
            for k in fe_f:
                assert fe_f[k] < 1.5 * fe_c[k] + 1e-3, (k, fe_f[k], fe_c[k])
            for k in ge_f:
                assert ge_f[k] < 1.5 * ge_c[k] + 2e-2, (k, ge_f[k], ge_c[k])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        fused_sa.ENABLED = True


# Operand-rounding emulation (oracle/cpu_modules.emulate_product_operands).  Measured on B200
# (profiles/r02/emulation_parity.log): one block -- features 1e-5 .. 2.6e-5, every gradient
# 1.5e-3 .. 8.8e-3; backbone -- features 1.7e-5 (sa1) / 1.9e-4 / 5.9e-4 / 1.5e-3 / 2.8e-3 (fp2).
# Why not 1e-6 everywhere although single layers agree to 7e-7 / 9e-6 (test_mlp_gpu.py): the next
# layer re-quantises its input to TF32, and a perturbation d of a value flips its rounding with
# probability d / 2^-10, each flip costing 2^-10 -- rms sqrt(d * 2^-10): 1e-7 -> 1e-5 -> 1e-4 ->
# the TF32 quantisation noise itself.  Gradients are discontinuous in the forward values (ReLU
# masks, max-pool winners): a fraction p of flipped decisions is a relative L2 difference of
# ~sqrt(2p), so 1e-5 forward differences cap gradient agreement at a few 1e-3 per block and at
# ~1e-1 through four stacked blocks, whatever the reference.
EMU_FEAT_TOL = {"sa1_features": 1e-4, "sa2_features": 6e-4, "sa3_features": 2e-3,
                "sa4_features": 5e-3, "fp2_features": 1e-2}
EMU_GRAD_TOL = 0.3
EMU_BLOCK_FEAT_TOL = 1e-4
EMU_BLOCK_GRAD_TOL = 2e-2


@pytest.mark.parametrize("npts", [8000, 40000])
def test_backbone_vs_operand_rounding_emulation(cuda, npts, capsys):
    """The PRODUCT path (fused tcgen05 SA blocks + dense tcgen05 FP layers) held to a tight bound at
    backbone scale: the oracle port with every 1x1 convolution's operands rounded exactly as the
    kernels round them (cvt.rna.tf32 forward; BF16 round-to-nearest-even in the fused SA backward;
    TF32 in the FP layers' backward; cpu_modules.emulate_product_operands) and exact products /
    sums (the pooled top layers with the backward's BF16 recompute of z: cpu_modules._TopConvBN).
    What is left is summation order -- amplified layer by layer by re-quantisation, see the
    comment above EMU_FEAT_TOL: sa1 agrees 30x tighter than with the fp32 reference, fp2 4x.
    Gradients through four stacked blocks stay at the network's own discontinuity floor."""
    from backtoreality_b200.backbone_module import Pointnet2Backbone
    from backtoreality_b200 import fused_sa
    from oracle import cpu_modules
    assert fused_sa.ENABLED
    torch.manual_seed(3)
    port = cpu_modules.emulate_product_operands(cpu_modules.Backbone(input_feature_dim=1).train())
    state = {k: v.clone() for k, v in port.state_dict().items()}
    pc = torch.from_numpy(scenes.batch(20, 2, npts, C=1, kind="room", dup=0.2))
    want = port(pc)
    (want["fp2_features"] * pattern_like(want["fp2_features"])).sum().backward()
    net = Pointnet2Backbone(input_feature_dim=1)
    net.load_state_dict(state)
    net = net.to(cuda).train()
    got = net(pc.to(cuda))
    for k in ("sa1_inds", "sa2_inds", "fp2_inds"):
        assert torch.equal(got[k].cpu(), want[k]), k
    fe = {k: rel_l2(got[k].detach().cpu().numpy(), want[k].detach().numpy())
          for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features")}
    (got["fp2_features"] * pattern_like(got["fp2_features"])).sum().backward()
    ge = {}
    for (n1, p1), (n2, p2) in zip(port.named_parameters(), net.named_parameters()):
        assert n1 == n2
        ge[n1] = rel_l2(p2.grad.cpu().numpy(), p1.grad.numpy())
    worst = sorted(ge.items(), key=lambda kv: -kv[1])[:4]
    with capsys.disabled():
        print("\n[emulation %d] features %s\n[emulation %d] gradients worst %s  median %.2e" % (
            npts, {k: "%.1e" % e for k, e in fe.items()}, npts,
            [(n, "%.1e" % e) for n, e in worst], float(np.median(list(ge.values())))))
    emu_log("backbone N=%d" % npts, grad_worst=worst[0][1],
            grad_median=float(np.median(list(ge.values()))), **fe)
    for k, e in fe.items():
        assert e < EMU_FEAT_TOL[k], (k, e)
    for k, e in ge.items():
        assert e < EMU_GRAD_TOL, (k, e)


@pytest.mark.parametrize("cfg", [dict(N=6000, C=1, npoint=512, radius=0.2, nsample=64, mlp=[1, 64, 64, 128]),
                                 dict(N=2048, C=128, npoint=1024, radius=0.4, nsample=32, mlp=[128, 128, 128, 256]),
                                 dict(N=1024, C=256, npoint=512, radius=0.8, nsample=16, mlp=[256, 128, 128, 256])])
def test_sa_block_vs_operand_rounding_emulation(cuda, cfg, capsys):
    """ONE fused SA block (SA1 / SA2 / SA3 shapes, train-mode BatchNorm), forward and every
    gradient, against the oracle port with the kernels' operand rounding: one block is shallow
    enough that re-quantisation has not yet amplified the fp32 differences, so this is the tight
    product-path bound (the fp32 comparison of the same block sits at 5e-4 / 3-4e-2)."""
    from backtoreality_b200.pointnet2_modules import PointnetSAModuleVotes
    from backtoreality_b200 import fused_sa
    from oracle import cpu_modules
    assert fused_sa.ENABLED
    torch.manual_seed(11)
    port = cpu_modules.emulate_product_operands(cpu_modules.SAModuleVotes(
        npoint=cfg["npoint"], radius=cfg["radius"], nsample=cfg["nsample"], mlp=list(cfg["mlp"]),
        normalize_xyz=True).train())
    sa = PointnetSAModuleVotes(npoint=cfg["npoint"], radius=cfg["radius"], nsample=cfg["nsample"],
                               mlp=list(cfg["mlp"]), use_xyz=True, normalize_xyz=True)
    sa.load_state_dict(port.state_dict())
    sa = sa.to(cuda).train()
    pc = torch.from_numpy(scenes.batch(31, 2, cfg["N"], C=0, kind="room", dup=0.2))[..., :3].contiguous()
    feats = torch.randn(2, cfg["C"], cfg["N"], generator=torch.Generator().manual_seed(5))
    f_cpu = feats.clone().requires_grad_(True)
    f_gpu = feats.to(cuda).requires_grad_(True)
    _, want, inds = port(pc, f_cpu)
    _, got, ginds = sa(pc.to(cuda), f_gpu)
    assert torch.equal(ginds.cpu(), inds)
    (want * pattern_like(want)).sum().backward()
    (got * pattern_like(got)).sum().backward()
    errs = {"features": rel_l2(got.detach().cpu().numpy(), want.detach().numpy()),
            "g_input": rel_l2(f_gpu.grad.cpu().numpy(), f_cpu.grad.numpy())}
    for (n1, p1), (n2, p2) in zip(port.named_parameters(), sa.named_parameters()):
        assert n1 == n2
        errs[n1.replace("mlp_module.", "")] = rel_l2(p2.grad.cpu().numpy(), p1.grad.numpy())
    emu_log("sa_block C=%d N=%d" % (cfg["C"], cfg["N"]), **errs)
    with capsys.disabled():
        print("\n[emulation block %s] %s" % (cfg["mlp"], {k: "%.1e" % e for k, e in errs.items()}))
    assert errs.pop("features") < EMU_BLOCK_FEAT_TOL
    for k, e in errs.items():
        assert e < EMU_BLOCK_GRAD_TOL, (k, e)


def test_captured_train_step_matches_eager(cuda):
    """The whole training step replayed from ONE CUDA graph (train_step.CapturedTrainStep) must
    compute what the eager step computes.  Run with a zero learning rate so the weights stay put:
    the forward is deterministic, so the replayed loss of every batch must equal the eager loss
    EXACTLY; gradients carry the run-to-run noise of float atomics amplified by the BatchNorm
    backward chain (measured 1e-3..7e-3 between two identical eager runs,
    scripts/bwd_determinism.py), hence a tolerance there."""
    from backtoreality_b200.train_step import CapturedTrainStep
    from backtoreality_b200.votenet import VoteNet

    def make():
        torch.manual_seed(5)
        net = VoteNet(4, 1, 4, np.ones((4, 3), np.float32), input_feature_dim=1, num_proposal=64,
                      vote_factor=1, sampling="vote_fps").to(cuda).train()
        opt = torch.optim.SGD(net.parameters(), lr=0.0)

        def step(pc):
            for p in net.parameters():
                p.grad = None
            ep = net({"point_clouds": pc})
            loss = (ep["proposal_scores_raw"] ** 2).mean() + ((ep["vote_xyz"] - ep["seed_xyz"]) ** 2).mean()
            loss.backward()
            opt.step()
            return loss.detach()
        return net, step

    batches = [torch.from_numpy(scenes.batch(300 + 2 * i, 2, 8192, C=1, kind="room", dup=0.2)).to(cuda)
               for i in range(4)]
    net_e, step_e = make()
    eager = [float(step_e(b)) for b in batches]
    net_g, step_g = make()
    captured = CapturedTrainStep(step_g, batches[0], warmup=3)
    assert captured.launches_per_step > 50   # libb2r launches inside the captured step
    got = [float(captured(b)) for b in batches]
    assert got == eager
    assert len(set(got)) == len(got)          # four different batches, four different losses
    for (n1, p1), (n2, p2) in zip(net_e.named_parameters(), net_g.named_parameters()):
        assert torch.equal(p1, p2), n1        # lr = 0
        assert rel_l2(p2.grad.cpu().numpy(), p1.grad.cpu().numpy()) < 3e-2, n1
    # BatchNorm bookkeeping advanced once per executed step: 3 warm-up + 4 replays vs 4 eager
    bn_e = net_e.backbone_net.sa1.mlp_module.layer0.bn.bn
    bn_g = net_g.backbone_net.sa1.mlp_module.layer0.bn.bn
    assert int(bn_e.num_batches_tracked) == 4 and int(bn_g.num_batches_tracked) == 7
    # ... unless the caller asks for the warm-up steps to be rolled back (snapshot=): the capture
    # then leaves weights, running statistics and step counts where they were
    net_s, step_s = make()
    before = {k: v.clone() for k, v in net_s.state_dict().items()}
    cap_s = CapturedTrainStep(step_s, batches[0], warmup=3, snapshot=[net_s])
    for k, v in net_s.state_dict().items():
        assert torch.equal(v, before[k]), k
    cap_s(batches[1])
    assert int(net_s.backbone_net.sa1.mlp_module.layer0.bn.bn.num_batches_tracked) == 1


def test_pipelined_train_step_at_the_benched_configuration(cuda):
    """bench.py's default arm exactly (BASELINE.json configs[1]): 8 scenes x 40000 points, the
    22-class VoteNet head with 256 proposals, PipelinedTrainStep with 4-CTA FPS clusters, the
    pre-pass started after SA level 2's forward and the default grid caps.  With lr = 0 call i must
    return the eager loss of batch i, and the rotated geometry must equal a fresh pre-pass."""
    from backtoreality_b200.train_step import PipelinedTrainStep
    from backtoreality_b200.votenet import VoteNet

    def make():
        torch.manual_seed(0)
        net = VoteNet(22, 1, 22, np.ones((22, 3), np.float32), input_feature_dim=1, num_proposal=256,
                      vote_factor=1, sampling="vote_fps").to(cuda).train()
        opt = torch.optim.SGD(net.parameters(), lr=0.0)

        def step(pc, geometry=None):
            for p in net.parameters():
                p.grad = None
            ep = net({"point_clouds": pc, "geometry": geometry})
            loss = (ep["proposal_scores_raw"] ** 2).mean() + ((ep["vote_xyz"] - ep["seed_xyz"]) ** 2).mean()
            loss.backward()
            opt.step()
            return loss.detach()
        return net, step

    batches = [torch.from_numpy(scenes.batch(1000 + 8 * i, 8, 40000, C=1, kind="room", dup=0.2)).to(cuda)
               for i in range(3)]
    net_e, step_e = make()
    eager = [float(step_e(b)) for b in batches]
    net_p, step_p = make()
    caps, head_cap = PipelinedTrainStep.default_caps(8, 4, 1)
    net_p.pnet.vote_aggregation.sm_limit = head_cap
    from backtoreality_b200.train_step import PipelinedTrainStepPP
    pipe = PipelinedTrainStepPP(net_p.backbone_net, step_p, batches[0], warmup=2, fps_cluster=4,
                                sm_caps=caps, start_after_level=1)
    got = [float(pipe(batches[(i + 1) % 3])) for i in range(3)]
    for g, e in zip(got, eager):
        assert abs(g - e) <= 2e-5 * abs(e), (got, eager)
    torch.cuda.synchronize()
    fresh = net_p.backbone_net.geometry_prepass(batches[0][..., :3].contiguous())
    torch.cuda.synchronize()
    for lv_p, lv_f in zip(pipe.geo_cur, fresh):
        for k in ("inds", "new_xyz", "idx"):
            assert torch.equal(lv_p[k], lv_f[k]), k


@pytest.mark.parametrize("start_after_level,pingpong", [(None, False), (1, False), (3, False),
                                                        (1, True), (None, True)])
def test_pipelined_train_step_matches_sequential(cuda, start_after_level, pingpong):
    """train_step.PipelinedTrainStep (geometry pre-pass of batch i+1 beside the step of batch i,
    started with the step or from the hook after an SA level's forward, narrow FPS clusters,
    capped MLP grids) must train on the same batches in the same order as
    the plain step: with lr = 0 the loss returned by call i is the eager loss of batch i (up to
    the summation order of the BatchNorm statistics under a different persistent grid), the
    rotated geometry buffers hold exactly the indices a fresh pre-pass computes, and the
    gradients agree within the atomics' run-to-run noise."""
    from backtoreality_b200.train_step import PipelinedTrainStep, PipelinedTrainStepPP
    from backtoreality_b200.votenet import VoteNet
    Pipe = PipelinedTrainStepPP if pingpong else PipelinedTrainStep   # two graphs, no rotation copies

    def make():
        torch.manual_seed(5)
        net = VoteNet(4, 1, 4, np.ones((4, 3), np.float32), input_feature_dim=1, num_proposal=64,
                      vote_factor=1, sampling="vote_fps").to(cuda).train()
        opt = torch.optim.SGD(net.parameters(), lr=0.0)

        def step(pc, geometry=None):
            for p in net.parameters():
                p.grad = None
            ep = net({"point_clouds": pc, "geometry": geometry})
            loss = (ep["proposal_scores_raw"] ** 2).mean() + ((ep["vote_xyz"] - ep["seed_xyz"]) ** 2).mean()
            loss.backward()
            opt.step()
            return loss.detach()
        return net, step

    batches = [torch.from_numpy(scenes.batch(300 + 2 * i, 2, 12000, C=1, kind="room", dup=0.2)).to(cuda)
               for i in range(4)]
    net_e, step_e = make()
    eager = [float(step_e(b)) for b in batches]
    net_p, step_p = make()
    net_p.pnet.vote_aggregation.sm_limit = (100, 120)
    pipe = Pipe(net_p.backbone_net, step_p, batches[0], warmup=2, fps_cluster=3,
                start_after_level=start_after_level)
    assert pipe.launches_per_step > 50
    got = [float(pipe(batches[(i + 1) % 4])) for i in range(4)]   # call i trains on batch i
    assert len(set(got)) == 4
    for g, e in zip(got, eager):
        assert abs(g - e) <= 1e-5 * abs(e), (got, eager)
    # after four calls the current batch is batches[0] again: its rotated geometry is the fresh one
    torch.cuda.synchronize()
    fresh = net_p.backbone_net.geometry_prepass(batches[0][..., :3].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(pipe.cur, batches[0])
    assert set(pipe.geo_cur[3]) == set(PipelinedTrainStep.GEO_KEYS + PipelinedTrainStep.FP_KEYS)
    for lv_p, lv_f in zip(pipe.geo_cur, fresh):
        for k in lv_p:
            n = None
            if k in ("cidx", "ccen"):      # a plan's buffers are only defined up to its total
                n = int(lv_f["cmeta"][8])
                assert n == int(lv_p["cmeta"][8])
            assert torch.equal(lv_p[k][:n], lv_f[k][:n]), k
    for (n1, p1), (n2, p2) in zip(net_e.named_parameters(), net_p.named_parameters()):
        assert torch.equal(p1, p2), n1        # lr = 0
        assert rel_l2(p2.grad.cpu().numpy(), p1.grad.cpu().numpy()) < 3e-2, n1


@pytest.mark.parametrize("start_after_level", [None, 1])
def test_pipelined_train_step_depth2_matches_sequential(cuda, start_after_level):
    """train_step.PipelinedTrainStep2: SA1's geometry of batch i+2 and the later levels' geometry
    of batch i+1 run beside the step of batch i.  Same contract as the depth-1 pipeline: call i
    returns the eager loss of batch i, the rotated buffers hold what a fresh pre-pass computes."""
    from backtoreality_b200.train_step import PipelinedTrainStep2
    from backtoreality_b200.votenet import VoteNet

    def make():
        torch.manual_seed(5)
        net = VoteNet(4, 1, 4, np.ones((4, 3), np.float32), input_feature_dim=1, num_proposal=64,
                      vote_factor=1, sampling="vote_fps").to(cuda).train()
        opt = torch.optim.SGD(net.parameters(), lr=0.0)

        def step(pc, geometry=None):
            for p in net.parameters():
                p.grad = None
            ep = net({"point_clouds": pc, "geometry": geometry})
            loss = (ep["proposal_scores_raw"] ** 2).mean() + ((ep["vote_xyz"] - ep["seed_xyz"]) ** 2).mean()
            loss.backward()
            opt.step()
            return loss.detach()
        return net, step

    batches = [torch.from_numpy(scenes.batch(300 + 2 * i, 2, 12000, C=1, kind="room", dup=0.2)).to(cuda)
               for i in range(5)]
    net_e, step_e = make()
    eager = [float(step_e(b)) for b in batches]
    net_p, step_p = make()
    net_p.pnet.vote_aggregation.sm_limit = (100, 120)
    pipe = PipelinedTrainStep2(net_p.backbone_net, step_p, batches[0], batches[1], warmup=2,
                               fps_cluster=3, start_after_level=start_after_level)
    pipe.prime(batches[0], batches[1])
    got = [float(pipe(batches[(i + 2) % 5])) for i in range(5)]   # call i trains on batch i
    assert len(set(got)) == 5
    for g, e in zip(got, eager):
        assert abs(g - e) <= 1e-5 * abs(e), (got, eager)
    # after five calls the current batch is batches[0] again and the next one batches[1]
    torch.cuda.synchronize()
    assert torch.equal(pipe.cur, batches[0]) and torch.equal(pipe.next, batches[1])
    fresh = net_p.backbone_net.geometry_prepass(batches[0][..., :3].contiguous())
    fresh1 = net_p.backbone_net.geometry_prepass(batches[1][..., :3].contiguous())
    torch.cuda.synchronize()
    for lv_p, lv_f in zip(pipe.geo_cur + [pipe.geo_a_next], fresh + [fresh1[0]]):
        for k in lv_p:
            n = int(lv_f["cmeta"][8]) if k in ("cidx", "ccen") else None
            assert torch.equal(lv_p[k][:n], lv_f[k][:n]), k


def test_geometry_stream_matches_serial_path(cuda):
    """Pointnet2Backbone with the geometry pre-pass on a side stream (FPS / centre gather / ball
    query of all levels ahead of the MLPs, MLP grids capped to leave SMs free) must compute what
    the serial path computes: identical indices and centres, features equal up to the summation
    order of the BatchNorm statistics (a different persistent grid), same gradients within the
    run-to-run noise of the atomics."""
    from backtoreality_b200 import backbone_module
    from backtoreality_b200.backbone_module import Pointnet2Backbone
    torch.manual_seed(21)
    net = Pointnet2Backbone(input_feature_dim=1).to(cuda).train()
    pc = torch.from_numpy(scenes.batch(400, 2, 20000, C=1, kind="room", dup=0.2)).to(cuda)
    outs = []
    old = backbone_module.GEOMETRY_STREAM
    try:
        for flag in (True, False):
            backbone_module.GEOMETRY_STREAM = flag
            for p in net.parameters():
                p.grad = None
            ep = net(pc)
            (ep["fp2_features"] * pattern_like(ep["fp2_features"])).sum().backward()
            torch.cuda.synchronize()
            outs.append(({k: v.detach().clone() for k, v in ep.items()},
                         {n: p.grad.clone() for n, p in net.named_parameters()}))
    finally:
        backbone_module.GEOMETRY_STREAM = old
    (ea, ga), (eb, gb) = outs
    for k in ("sa1_inds", "sa2_inds", "fp2_inds", "sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz"):
        assert torch.equal(ea[k], eb[k]), k
    for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
        assert rel_l2(ea[k].cpu().numpy(), eb[k].cpu().numpy()) < 1e-5, k
    for n in ga:
        assert rel_l2(ga[n].cpu().numpy(), gb[n].cpu().numpy()) < 3e-2, n


def test_host_prefetcher_delivers_batches_in_order(cuda):
    """train_step.HostPrefetcher: uploads run on a copy stream one step ahead; what the compute
    stream reads from a slot is exactly the batch uploaded into it, also when the slot is reused
    while earlier consumers are still queued."""
    from backtoreality_b200.train_step import HostPrefetcher
    host = [torch.full((4, 50000, 4), float(i)).pin_memory() for i in range(6)]
    pre = HostPrefetcher(host[0], cuda)
    sums = []
    slot = pre.upload(host[0])
    for i in range(6):
        nslot = pre.upload(host[(i + 1) % 6])
        b = pre.get(slot)
        # slow consumer; min == max == i proves the slot holds ONE whole batch
        sums.append(torch.stack([b.min(), b.max()]) + 0 * torch.randn(1 << 20, device=cuda).sum())
        pre.release(slot)
        slot = nslot
    torch.cuda.synchronize()
    assert [s.tolist() for s in sums] == [[float(i), float(i)] for i in range(6)]
    assert pre.bytes_uploaded == 7 * host[0].numel() * 4


def test_prepacked_weight_images_match_inline_packing(cuda):
    """fused_sa.prepack (operand images packed on a side stream at the start of the step) must
    hand the layers the same images inline packing produces, be consumed once, and notice a
    weight that changed after packing."""
    from backtoreality_b200 import fused_sa
    from backtoreality_b200.pointnet2_modules import PointnetSAModuleVotes
    torch.manual_seed(3)
    sa = PointnetSAModuleVotes(npoint=128, radius=0.4, nsample=32, mlp=[16, 32, 32, 64],
                               use_xyz=True, normalize_xyz=True).to(cuda).train()
    fused_sa.prepack([sa.mlp_module])
    fused_sa.prepack_join()
    torch.cuda.synchronize()
    assert len(fused_sa._PREPACKED) == 6
    for i, blk in enumerate(sa.mlp_module):
        w = blk.conv.weight
        a = fused_sa._take(w, "tf32")
        b = fused_sa._take(w, "bf16")
        torch.cuda.synchronize()
        assert torch.equal(a, fused_sa.pack_weight(w, gather=(i == 0)))
        assert torch.equal(b.view(torch.int16), fused_sa.pack_weight_bf16(w, gather=(i == 0)).view(torch.int16))
        assert fused_sa._take(w, "tf32") is None          # consumed
    fused_sa.prepack([sa.mlp_module])
    with torch.no_grad():
        sa.mlp_module[0].conv.weight.mul_(2.0)            # stale image must not be used
    assert fused_sa._take(sa.mlp_module[0].conv.weight, "tf32") is None
    assert fused_sa._take(sa.mlp_module[1].conv.weight, "tf32") is not None
    fused_sa.prepack_join()
    # end to end: the block computes the same output with and without pre-packed images
    xyz = torch.rand(2, 1000, 3, device=cuda)
    feats = torch.randn(2, 16, 1000, device=cuda)
    fused_sa.prepack([])
    _, y0, _ = sa(xyz, feats)
    fused_sa.prepack([sa.mlp_module])
    feats.requires_grad_(True)
    _, y1, _ = sa(xyz, feats)
    assert sorted(k[1] for k in fused_sa._PREPACKED) == ["bf16"] * 3   # forward took the TF32 images
    y1.square().mean().backward()
    assert len(fused_sa._PREPACKED) == 0                               # backward took the BF16 ones
    fused_sa.prepack_join()
    assert torch.equal(y0, y1)
    # a different module must never be served this module's images (entries are keyed by the
    # Parameter object, not by its address)
    fused_sa.prepack([sa.mlp_module])
    other = PointnetSAModuleVotes(npoint=128, radius=0.4, nsample=32, mlp=[16, 32, 32, 64],
                                  use_xyz=True, normalize_xyz=True).to(cuda).train()
    assert fused_sa._take(other.mlp_module[0].conv.weight, "tf32") is None
    fused_sa.prepack([])


def test_two_forwards_one_backward_is_the_sum_of_two_steps(cuda):
    """The BR training step (BASELINE.json configs[2], reference train_Votenet_BR.py:277-289) runs
    TWO forwards (source and target batch) through the same network and ONE backward of the summed
    loss.  Gradients are linear in the loss, and train-mode BatchNorm makes each forward
    independent of the other, so grad(L_a + L_b) must equal grad(L_a) + grad(L_b) of two separate
    steps (within the atomics' run-to-run noise); every BatchNorm sees two batches."""
    from backtoreality_b200.votenet import VoteNet
    torch.manual_seed(9)
    net = VoteNet(4, 1, 4, np.ones((4, 3), np.float32), input_feature_dim=1, num_proposal=64,
                  vote_factor=1, sampling="vote_fps").to(cuda).train()
    pa, pb = (torch.from_numpy(scenes.batch(500 + 2 * i, 2, 9000, C=1, kind="room", dup=0.2)).to(cuda)
              for i in range(2))

    def loss_of(pc):
        ep = net({"point_clouds": pc})
        return (ep["proposal_scores_raw"] ** 2).mean() + ((ep["vote_xyz"] - ep["seed_xyz"]) ** 2).mean()

    def grads():
        g = {n: p.grad.clone() for n, p in net.named_parameters()}
        for p in net.parameters():
            p.grad = None
        return g

    bn = net.backbone_net.sa2.mlp_module.layer1.bn.bn
    n0 = int(bn.num_batches_tracked)
    (loss_of(pa) + loss_of(pb)).backward()          # the BR pattern
    assert int(bn.num_batches_tracked) == n0 + 2
    both = grads()
    loss_of(pa).backward()
    ga = grads()
    loss_of(pb).backward()
    gb = grads()
    for n in both:
        want = (ga[n] + gb[n]).cpu().numpy()
        assert rel_l2(both[n].cpu().numpy(), want) < 3e-2, n


# ---------------- goldens from the reference's own Python (make_golden.py) ----
def test_centers_offset_and_three_nn_interpolate_golden(cuda, fp32_mlp):
    """PointnetSAModuleCenters (V pointnet2_modules.py:357-451), GroupFree3D's
    PointnetSAModuleOffset (:481-576) and ThreeNNInterpolate (:722-730) against fixtures generated
    from the reference's unmodified modules."""
    from backtoreality_b200.pointnet2_modules import (PointnetSAModuleCenters, PointnetSAModuleOffset,
                                                      ThreeNNInterpolate)
    g = golden("centers_offset.npz")
    ftol, gtol = fp32_mlp
    seed = int(g["seed"])
    torch.manual_seed(seed)
    head = PointnetSAModuleCenters(npoint=32, radius=0.8, nsample=16, mlp=[64, 32], use_xyz=True,
                                   normalize_xyz=False)
    assert abs(weight_checksum(head) - float(g["wsum"])) < 1e-6 * float(g["wsum"])
    head = head.to(cuda)
    gen = torch.Generator().manual_seed(seed + 1)
    xyz0 = torch.rand(2, 300, 3, generator=gen) * 3.0
    feats0 = torch.randn(2, 64, 300, generator=gen)
    centers0 = xyz0[:, :32] + 0.05 * torch.randn(2, 32, 3, generator=gen)
    xyz = xyz0.to(cuda).requires_grad_(True)
    feats = feats0.to(cuda).requires_grad_(True)
    centers = centers0.to(cuda).requires_grad_(True)
    y = head(xyz, feats, centers)
    assert rel_l2(y.detach().cpu().numpy(), g["y"]) < ftol
    (y * pattern_like(y)).sum().backward()
    assert rel_l2(xyz.grad.cpu().numpy(), g["g_xyz"]) < gtol
    assert rel_l2(centers.grad.cpu().numpy(), g["g_centers"]) < gtol
    assert rel_l2(sub(feats.grad), g["g_feats"]) < gtol
    assert rel_l2(head.mlp_module.layer0.conv.weight.grad.cpu().numpy(), g["g_w"]) < gtol
    torch.manual_seed(seed)
    off = PointnetSAModuleOffset(npoint=32, radius=0.6, nsample=16, mlp=[64, 32, 32, 48],
                                 use_xyz=True, normalize_xyz=True)
    assert abs(weight_checksum(off) - float(g["wsum_off"])) < 1e-6 * float(g["wsum_off"])
    y2 = off.to(cuda)(xyz.detach(), feats.detach(), centers.detach())
    assert rel_l2(y2.detach().cpu().numpy(), g["y_off"]) < ftol
    tni = ThreeNNInterpolate(feats.detach()[:, :, :40].contiguous(), xyz.detach()[:, :40].contiguous(),
                             xyz.detach()[:, 100:180].contiguous())
    assert rel_l2(tni.cpu().numpy(), g["tni"]) < 1e-6


def test_jitter_backbone_golden(cuda):
    """Pointnet2Backbone_jitter (V models/backbone_module.py:136-262) with centres, product path
    (fused SA blocks + dense FP layers) against the reference's fp32 CPU run."""
    from backtoreality_b200.backbone_module import Pointnet2Backbone_jitter
    g = golden("backbone_jitter.npz")
    torch.manual_seed(int(g["seed"]))
    net = Pointnet2Backbone_jitter(input_feature_dim=1)
    assert abs(weight_checksum(net) - float(g["wsum"])) < 1e-6 * float(g["wsum"])
    net = net.to(cuda).train()
    pc = torch.from_numpy(scenes.batch(60, 2, 3000, C=1, kind="room", dup=0.2)).to(cuda)
    cx = torch.from_numpy(g["center_xyz"]).to(cuda)
    cc = torch.from_numpy(g["center_cls"]).to(cuda)
    ep = net(pc, cx, cc)
    cf = ep["center_features"]
    assert cf.shape == (2, 128 + 22, 64)
    assert torch.equal(cf[:, 128:].detach().cpu(), torch.from_numpy(g["center_features"][:, 128:]))
    # five TF32 blocks + two dense FP modules deep
    assert rel_l2(sub(ep["fp2_features"]), g["fp2_features"]) < 2e-2
    assert rel_l2(cf.detach().cpu().numpy(), g["center_features"]) < 2e-2
    (cf * pattern_like(cf)).sum().backward()
    # gradients through stacked TF32/BF16 blocks: compared by direction and magnitude (see
    # profiles/r01/tf32_gradient_noise_cudnn_vs_fused.log for the cuDNN-TF32 noise floor)
    for name, got in (("g_ctjt", net.ctjt_head.mlp_module.layer0.conv.weight.grad.cpu().numpy()),
                      ("g_fp2_l1", sub(net.fp2.mlp.layer1.conv.weight.grad)),
                      ("g_sa1_l0", sub(net.sa1.mlp_module.layer0.conv.weight.grad))):
        want = g[name]
        cos = float((got.ravel() * want.ravel()).sum() /
                    (np.linalg.norm(got) * np.linalg.norm(want) + 1e-30))
        assert cos > 0.95, (name, cos)
        assert rel_l2(got, want) < 0.35, name


@pytest.mark.parametrize("train", [True, False])
def test_vote_heads_golden(cuda, train):
    """VotingModule -> L2 normalisation -> ProposalModule (seed_fps) against fixtures from the
    reference's voting_module.py:38-65 / votenet.py:93-94 / proposal_module.py:52-120: the dense
    tcgen05 heads around the fused vote aggregation."""
    from backtoreality_b200 import dense_mlp
    from backtoreality_b200.votenet import ProposalModule, VotingModule
    g = golden("vote_heads_train.npz" if train else "vote_heads_eval.npz")
    seed = int(g["seed"])
    torch.manual_seed(seed)
    vgen = VotingModule(1, 256)
    pnet = ProposalModule(4, 2, 4, g["msa"], 32, "seed_fps")
    assert abs(weight_checksum(vgen) + weight_checksum(pnet) - float(g["wsum"])) < 1e-6 * float(g["wsum"])
    for m in list(vgen.modules()) + list(pnet.modules()):
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.momentum = 0.2
            if not train:
                m.running_mean.data = torch.randn(m.running_mean.shape) * 0.1
                m.running_var.data = torch.rand(m.running_var.shape) + 0.5
    vgen = vgen.to(cuda).train(train)
    pnet = pnet.to(cuda).train(train)
    gen = torch.Generator().manual_seed(seed + 1)
    seed_xyz = (torch.rand(2, 256, 3, generator=gen) * 3.0).to(cuda).requires_grad_(True)
    seed_feat = torch.randn(2, 256, 256, generator=gen).to(cuda).requires_grad_(True)
    assert dense_mlp.enabled()
    res = vgen.forward_normalized(seed_xyz, seed_feat)
    assert res is not None
    vote_xyz, vote_feat, vote_feat_pm = res
    ep = pnet(vote_xyz, vote_feat, {"seed_xyz": seed_xyz}, features_pm=vote_feat_pm)
    assert np.array_equal(ep["aggregated_vote_inds"].cpu().numpy(), g["inds"])
    tol = 1e-2
    assert rel_l2(vote_xyz.detach().cpu().numpy(), g["vote_xyz"]) < tol
    assert rel_l2(sub(vote_feat), g["vote_feat"]) < tol
    assert rel_l2(ep["aggregated_vote_xyz"].detach().cpu().numpy(), g["agg_xyz"]) < tol
    assert rel_l2(sub(ep["aggregated_vote_features"]), g["agg_feat"]) < 3e-2
    assert rel_l2(ep["objectness_scores"].detach().cpu().numpy(), g["objectness"]) < 5e-2
    assert rel_l2(ep["center"].detach().cpu().numpy(), g["center"]) < 2e-2
    assert rel_l2(sub(ep["size_residuals"]), g["size_residuals"]) < 5e-2
    assert rel_l2(ep["sem_cls_scores"].detach().cpu().numpy(), g["sem_cls"]) < 5e-2
    loss = ((ep["objectness_scores"] * 0.7).sum() + (ep["center"] * 0.3).sum() +
            (ep["size_residuals"] * pattern_like(ep["size_residuals"])).sum() +
            (ep["sem_cls_scores"] * pattern_like(ep["sem_cls_scores"])).sum() + (vote_xyz * 0.11).sum())
    loss.backward()
    got = {"g_seed_xyz": seed_xyz.grad.cpu().numpy(), "g_seed_feat": sub(seed_feat.grad),
           "g_vgen_c1": sub(vgen.conv1.weight.grad), "g_vgen_c3": sub(vgen.conv3.weight.grad),
           "g_vgen_c3_b": vgen.conv3.bias.grad.cpu().numpy(), "g_vgen_bn2": vgen.bn2.weight.grad.cpu().numpy(),
           "g_pnet_c1": sub(pnet.conv1.weight.grad), "g_pnet_c3_b": pnet.conv3.bias.grad.cpu().numpy()}
    for name, v in got.items():
        want = g[name]
        cos = float((v.ravel() * want.ravel()).sum() / (np.linalg.norm(v) * np.linalg.norm(want) + 1e-30))
        assert cos > 0.97 and rel_l2(v, want) < 0.25, (name, cos, rel_l2(v, want))
    if train:
        assert rel_l2(vgen.bn1.running_mean.cpu().numpy(), g["rm_vgen_bn1"]) < 1e-2
        assert rel_l2(pnet.bn2.running_var.cpu().numpy(), g["rv_pnet_bn2"]) < 3e-2
        assert float(pnet.conv2.bias.grad.abs().max()) < 1e-3      # bias before a training-mode BN
    else:
        assert rel_l2(pnet.conv2.bias.grad.cpu().numpy(), g["g_pnet_c2_b"]) < 0.25


def test_flat_adam_matches_torch_adam(cuda):
    """flat_adam.FlatAdam (one kernel over the flat parameter buffer, device-side step counter)
    against torch.optim.Adam on the same gradients, with and without weight decay, and replayed
    from a CUDA graph."""
    from backtoreality_b200.flat_adam import FlatAdam
    for wd in (0.0, 0.01):
        torch.manual_seed(3)
        shapes = [(64, 4, 1, 1), (64,), (128, 131, 1, 1), (7,), (3, 5)]
        ref = [torch.nn.Parameter(torch.randn(s, device=cuda)) for s in shapes]
        ours = [torch.nn.Parameter(p.detach().clone()) for p in ref]
        o_ref = torch.optim.Adam(ref, lr=1e-2, weight_decay=wd)
        o_b2r = FlatAdam(ours, lr=1e-2, weight_decay=wd)
        for it in range(6):
            grads = [torch.randn_like(p) * (it + 1) for p in ref]
            for p, g in zip(ref, grads):
                p.grad = g.clone()
            o_ref.step()
            o_b2r.step(grads)
            for a, b in zip(ref, ours):
                torch.testing.assert_close(b, a, rtol=2e-5, atol=2e-6)
        # parameters are views of the flat buffer and their version counters moved
        assert all(p.data_ptr() >= o_b2r.flat_p.data_ptr() for p in ours)
        v0 = ours[0]._version
        o_b2r.step([torch.zeros_like(p) for p in ours])
        assert ours[0]._version > v0
    # graph replay: the step counter advances on the device
    p_g = [torch.nn.Parameter(torch.ones(1000, device=cuda))]
    p_e = [torch.nn.Parameter(torch.ones(1000, device=cuda))]
    og, oe = FlatAdam(p_g, lr=1e-2), torch.optim.Adam(p_e, lr=1e-2)
    gstat = torch.full((1000,), 0.5, device=cuda)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        og.step([gstat])
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        og.step([gstat])
    for _ in range(4):
        g.replay()
    for _ in range(6):      # 1 eager + 1 capture-time (not executed) ... count the executed ones
        p_e[0].grad = gstat.clone()
        oe.step()
    torch.cuda.synchronize()
    # warm-up step + 4 replays = 5 executed steps (the capture itself does not execute)
    p_chk = [torch.nn.Parameter(torch.ones(1000, device=cuda))]
    oc = torch.optim.Adam(p_chk, lr=1e-2)
    for _ in range(5):
        p_chk[0].grad = gstat.clone()
        oc.step()
    torch.testing.assert_close(p_g[0], p_chk[0], rtol=2e-5, atol=2e-6)


def test_gf3d_query_sampling_golden(cuda):
    """GroupFree3D query sampling against fixtures from the reference's
    G/models/modules.py:16-100: KPS objectness head and learned position embedding on the dense
    tcgen05 path (TF32), FPS / general sampling on libb2r's FPS + gather, gradients included."""
    from backtoreality_b200 import gf3d_modules as m
    g = golden("gf3d_query_sampling.npz")
    seed = int(g["seed"])
    torch.manual_seed(seed)
    cls_head = m.PointsObjClsModule(288)
    pos = m.PositionEmbeddingLearned(3, 288)
    pos6 = m.PositionEmbeddingLearned(6, 288)
    fps, gs = m.FPSModule(64), m.GeneralSamplingModule()
    wsum = weight_checksum(cls_head) + weight_checksum(pos) + weight_checksum(pos6)
    assert abs(wsum - float(g["wsum"])) < 1e-6 * float(g["wsum"])    # same init order as the reference
    for mod in list(cls_head.modules()) + list(pos.modules()) + list(pos6.modules()):
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.momentum = 0.2
    cls_head, pos, pos6 = cls_head.to(cuda).train(), pos.to(cuda).train(), pos6.to(cuda).train()
    gen = torch.Generator().manual_seed(seed + 1)
    xyz = (torch.rand(2, 256, 3, generator=gen) * 3.0).to(cuda)
    feat = torch.randn(2, 288, 256, generator=gen).to(cuda).requires_grad_(True)
    logits = cls_head(feat)
    assert rel_l2(logits.detach().cpu().numpy(), g["logits"]) < 1e-2
    # the top-k of near-equal TF32 scores may order differently: sample with the reference's indices
    kps_inds = torch.from_numpy(g["kps_inds"]).to(cuda)
    k_xyz, k_feat, _ = gs(xyz, feat, kps_inds)
    f_xyz, f_feat, f_inds = fps(xyz, feat)
    assert np.array_equal(f_inds.cpu().numpy(), g["f_inds"])
    assert np.array_equal(k_xyz.cpu().numpy(), g["k_xyz"]) and np.array_equal(f_xyz.cpu().numpy(), g["f_xyz"])
    assert np.array_equal(sub(k_feat), g["k_feat"]) and np.array_equal(sub(f_feat), g["f_feat"])
    emb = pos(f_xyz)
    emb6 = pos6(torch.cat([f_xyz, f_xyz * 0.5 + 0.1], -1))
    assert rel_l2(sub(emb), g["emb"]) < 1e-2 and rel_l2(sub(emb6), g["emb6"]) < 1e-2
    loss = ((logits * pattern_like(logits)).sum() + (k_feat * 0.3).sum() + (f_feat * pattern_like(f_feat)).sum() +
            (emb * pattern_like(emb)).sum() + (emb6 * 0.01).sum())
    loss.backward()
    assert rel_l2(sub(feat.grad), g["g_feat"]) < 3e-2
    assert rel_l2(sub(cls_head.conv1.weight.grad), g["g_cls_c1"]) < 5e-2
    assert rel_l2(cls_head.conv3.bias.grad.cpu().numpy(), g["g_cls_c3_b"]) < 1e-3
    assert rel_l2(pos.position_embedding_head[0].weight.grad.cpu().numpy(), g["g_pos_c0"]) < 5e-2
    assert rel_l2(sub(pos.position_embedding_head[3].weight.grad), g["g_pos_c3"]) < 5e-2
    assert rel_l2(cls_head.bn1.running_mean.cpu().numpy(), g["rm_cls_bn1"]) < 1e-2
    # detector.py:150-175 restated: both branches fill the reference's end_points keys
    ep = {"fp2_xyz": xyz, "fp2_features": feat.detach(), "fp2_inds": torch.arange(256, device=cuda).repeat(2, 1)}
    cx, cf = m.sample_queries(ep, "fps", 64, fps_module=fps)
    assert torch.equal(ep["query_points_sample_inds"], f_inds) and cx.shape == (2, 64, 3) and cf.shape == (2, 288, 64)
    cx, cf = m.sample_queries(ep, "kps", 64, points_obj_cls=cls_head, gsample_module=gs)
    assert ep["seeds_obj_cls_logits"].shape == (2, 1, 256) and ep["query_points_xyz"].shape == (2, 64, 3)
    proj = torch.nn.Conv1d(288, 288, 1).to(cuda)
    want = proj(feat.detach())
    assert rel_l2(m.project_pm(proj, feat.detach()).detach().cpu().numpy(), want.detach().cpu().numpy()) < 5e-3


def test_split_geometry_equals_separate_prepasses(cuda):
    """geometry_prepass(plan_splits=2) + split_geometry (the BR step: one pre-pass over source +
    target scenes, consumed as two forwards) gives each half exactly the geometry -- indices AND
    pad-free plans -- of a pre-pass over that half alone, and the same forward output."""
    from backtoreality_b200.backbone_module import Pointnet2Backbone
    torch.manual_seed(2)
    net = Pointnet2Backbone(input_feature_dim=1).to(cuda).train()
    pc = torch.from_numpy(scenes.batch(500, 4, 12000, C=1, kind="room", dup=0.2)).to(cuda)
    xyz = pc[..., :3].contiguous()
    both = net.geometry_prepass(xyz, plan_splits=2)
    torch.cuda.synchronize()
    halves = Pointnet2Backbone.split_geometry(both, 2)
    for i, part in enumerate((pc[:2], pc[2:])):
        alone = net.geometry_prepass(part[..., :3].contiguous())
        torch.cuda.synchronize()
        for lv_s, lv_a in zip(halves[i], alone):
            for k in ("inds", "new_xyz", "idx"):
                assert torch.equal(lv_s[k], lv_a[k]), (i, k)
            if "cmeta" in lv_a:
                n = int(lv_a["cmeta"][8])
                assert n == int(lv_s["cmeta"][8])
                assert torch.equal(lv_s["cidx"][:n], lv_a["cidx"][:n]) and torch.equal(lv_s["ccen"][:n], lv_a["ccen"][:n])
        for k in ("fp1_idx", "fp1_weight", "fp2_idx", "fp2_weight"):
            assert torch.equal(halves[i][3][k], alone[3][k]), k
        with torch.no_grad():
            a = net(part, geometry=halves[i])["fp2_features"]
            b = net(part, geometry=alone)["fp2_features"]
        assert torch.equal(a, b)
