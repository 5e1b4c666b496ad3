"""The oracle's module-level port (oracle/cpu_modules.py) against the golden fixtures that were
generated from the reference's own Python stack (tests/golden/make_golden.py).  No GPU."""
import numpy as np
import pytest
import torch

from _util import golden, pattern_like, rel_l2, sub, weight_checksum
from backtoreality_b200 import scenes
from oracle import cpu_modules

TOL = 1e-5  # same torch CPU fp32 kernels, same op order: differences are thread-sum noise only


@pytest.mark.parametrize("fixture", ["backbone_votenet_eval.npz", "backbone_votenet_train.npz",
                                     "backbone_gf3d_train.npz"])
def test_backbone_port_matches_reference_python(fixture):
    g = golden(fixture)
    torch.manual_seed(int(g["seed"]))
    net = cpu_modules.Backbone(input_feature_dim=int(g["C"]), fp2_out=int(g["fp2_out"]))
    assert abs(weight_checksum(net) - float(g["wsum"])) < 1e-6 * float(g["wsum"])
    net.train(bool(g["train"]))
    pc = torch.from_numpy(scenes.batch(50, int(g["B"]), int(g["N"]), C=int(g["C"]), kind="room",
                                       dup=0.2))
    ep = net(pc)
    assert np.array_equal(ep["sa1_inds"].numpy(), g["sa1_inds"])
    assert np.array_equal(ep["sa2_inds"].numpy(), g["sa2_inds"])
    for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
        assert rel_l2(sub(ep[k]), g[k]) < TOL, k
    (ep["fp2_features"] * pattern_like(ep["fp2_features"])).sum().backward()
    assert rel_l2(sub(net.sa1.mlp_module.layer0.conv.weight.grad), g["g_sa1_l0"]) < 1e-4
    assert rel_l2(sub(net.sa2.mlp_module.layer0.conv.weight.grad), g["g_sa2_l0"]) < 1e-4
    assert rel_l2(sub(net.fp1.mlp.layer0.conv.weight.grad), g["g_fp1_l0"]) < 1e-4
    if g["train"]:
        bn = net.sa1.mlp_module.layer0.bn.bn
        assert rel_l2(bn.running_mean.numpy(), g["rm_sa1_l0"]) < TOL
        assert rel_l2(bn.running_var.numpy(), g["rv_sa1_l0"]) < TOL


def test_vote_aggregation_port():
    g = golden("vote_aggregation.npz")
    torch.manual_seed(int(g["seed"]))
    sa = cpu_modules.SAModuleVotes(npoint=64, radius=0.3, nsample=16, mlp=[32, 32, 32, 32],
                                   normalize_xyz=True)
    assert abs(weight_checksum(sa) - float(g["wsum"])) < 1e-6 * float(g["wsum"])
    gen = torch.Generator().manual_seed(int(g["seed"]) + 1)
    xyz = (torch.rand(2, 256, 3, generator=gen) * 2.0 + 0.5).requires_grad_(True)
    feats = torch.randn(2, 32, 256, generator=gen).requires_grad_(True)
    new_xyz, new_feats, inds = sa(xyz, feats)
    assert np.array_equal(inds.numpy(), g["inds"])
    assert np.array_equal(new_xyz.detach().numpy(), g["new_xyz"])
    assert rel_l2(new_feats.detach().numpy(), g["new_feats"]) < TOL
    ((new_feats * pattern_like(new_feats)).sum() + (new_xyz * 0.37).sum()).backward()
    assert rel_l2(xyz.grad.numpy(), g["g_xyz"]) < 1e-4
    assert rel_l2(sub(feats.grad), g["g_feats"]) < 1e-4


def test_fp_module_port():
    g = golden("fp_module.npz")
    torch.manual_seed(int(g["seed"]))
    fp = cpu_modules.FPModule(mlp=[64, 32, 24])
    assert abs(weight_checksum(fp) - float(g["wsum"])) < 1e-6 * float(g["wsum"])
    gen = torch.Generator().manual_seed(int(g["seed"]) + 1)
    unknown = torch.rand(2, 100, 3, generator=gen)
    known = torch.rand(2, 37, 3, generator=gen)
    known[:, 5] = known[:, 2]
    uf = torch.randn(2, 16, 100, generator=gen).requires_grad_(True)
    kf = torch.randn(2, 48, 37, generator=gen).requires_grad_(True)
    y = fp(unknown, known, uf, kf)
    assert rel_l2(y.detach().numpy(), g["y"]) < TOL
    (y * pattern_like(y)).sum().backward()
    assert rel_l2(kf.grad.numpy(), g["g_kf"]) < 1e-4
    assert rel_l2(sub(uf.grad), g["g_uf"]) < 1e-4


def test_new_module_parameter_layouts_match_the_reference_checksums():
    """the product's restatements of PointnetSAModuleCenters / Offset,
    Pointnet2Backbone_jitter, VotingModule and ProposalModule create their parameters in the
    reference's order and shapes: seeded initialisation reproduces the weight checksums stored in
    the fixtures that tests/golden/make_golden.py generated from the reference's own modules."""
    import numpy as np
    import torch
    from _util import golden, weight_checksum
    from backtoreality_b200.backbone_module import Pointnet2Backbone_jitter
    from backtoreality_b200.pointnet2_modules import PointnetSAModuleCenters, PointnetSAModuleOffset
    from backtoreality_b200.votenet import ProposalModule, VotingModule
    g = golden("centers_offset.npz")
    torch.manual_seed(int(g["seed"]))
    head = PointnetSAModuleCenters(npoint=32, radius=0.8, nsample=16, mlp=[64, 32], use_xyz=True,
                                   normalize_xyz=False)
    assert abs(weight_checksum(head) - float(g["wsum"])) < 1e-6 * float(g["wsum"])
    torch.manual_seed(int(g["seed"]))
    off = PointnetSAModuleOffset(npoint=32, radius=0.6, nsample=16, mlp=[64, 32, 32, 48],
                                 use_xyz=True, normalize_xyz=True)
    assert abs(weight_checksum(off) - float(g["wsum_off"])) < 1e-6 * float(g["wsum_off"])
    g = golden("backbone_jitter.npz")
    torch.manual_seed(int(g["seed"]))
    net = Pointnet2Backbone_jitter(input_feature_dim=1)
    assert abs(weight_checksum(net) - float(g["wsum"])) < 1e-6 * float(g["wsum"])
    assert "ctjt_head.mlp_module.layer0.conv.weight" in net.state_dict()
    g = golden("vote_heads_train.npz")
    torch.manual_seed(int(g["seed"]))
    vgen = VotingModule(1, 256)
    pnet = ProposalModule(4, 2, 4, g["msa"], 32, "seed_fps")
    assert abs(weight_checksum(vgen) + weight_checksum(pnet) - float(g["wsum"])) < 1e-6 * float(g["wsum"])
    assert pnet.conv3.out_channels == 2 + 3 + 2 * 2 + 4 * 4 + 4


def test_operand_rounding_emulation_of_the_port():
    """cpu_modules.emulate_product_operands (the checker of the GPU emulation tests): rounding
    helpers are bit-exact models of cvt.rna.tf32 / bf16 RNE; with "fp32" operands the emulated
    convolutions and the fused top conv + BatchNorm function reproduce the plain port (forward,
    every gradient, running statistics); state-dict keys do not change; switching off restores
    the plain modules."""
    import struct
    from backtoreality_b200 import scenes

    def f32(bits):
        return struct.unpack("<f", struct.pack("<I", bits))[0]

    x = torch.tensor([f32(0x3F800FFF), f32(0x3F801000), f32(0x3F801001), f32(0xBF801000), 0.0])
    assert [hex(v) for v in cpu_modules.round_tf32(x).view(torch.int32).tolist()] == \
        [hex(v) for v in torch.tensor([0x3F800000, 0x3F802000, 0x3F802000, 0xBF802000 - (1 << 32), 0],
                                      dtype=torch.int64).to(torch.int32).tolist()]   # ties away from zero
    y = torch.tensor([f32(0x3F808000), f32(0x3F818000), f32(0x3F808001)])
    assert cpu_modules.round_bf16(y).view(torch.int32).tolist() == [0x3F800000, 0x3F820000, 0x3F810000]

    torch.manual_seed(11)
    port = cpu_modules.SAModuleVotes(npoint=128, radius=0.4, nsample=16, mlp=[8, 16, 16, 32],
                                     normalize_xyz=True).train()
    keys = list(port.state_dict().keys())
    pc = torch.from_numpy(scenes.batch(31, 2, 1024, C=0))[..., :3].contiguous()
    feats = torch.randn(2, 8, 1024)
    w = None

    def run():
        nonlocal w
        port.zero_grad()
        port.mlp_module.layer2.bn.bn.reset_running_stats()
        f = feats.clone().requires_grad_(True)
        _, out, _ = port(pc, f)
        if w is None:
            w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
        (out * w).sum().backward()
        g = {n: p.grad.clone() for n, p in port.named_parameters()}
        g["input"] = f.grad.clone()
        return out.detach().clone(), g, port.mlp_module.layer2.bn.bn.running_var.clone()

    y0, g0, rv0 = run()
    cpu_modules.emulate_product_operands(port)
    assert list(port.state_dict().keys()) == keys
    assert port.mlp_module.layer0.conv.operands == ("tf32", "bf16")
    y_e, g_e, _ = run()
    assert 1e-5 < rel_l2(y_e, y0) < 5e-3            # TF32 operands are visible ...
    for m in port.modules():
        if isinstance(m, cpu_modules._Conv1x1):
            m.operands = ("fp32", "fp32")
    y1, g1, rv1 = run()                             # ... exact operands are not
    assert rel_l2(y1, y0) < 1e-6 and rel_l2(rv1, rv0) < 1e-6
    for n in g0:
        assert rel_l2(g1[n], g0[n]) < 2e-5, n
    cpu_modules.emulate_product_operands(port, False)
    y2, g2, _ = run()
    assert torch.equal(y2, y0) and all(torch.equal(g2[n], g0[n]) for n in g0)
