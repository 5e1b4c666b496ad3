"""Generate the golden fixtures in this directory from the REFERENCE ITSELF.

Runs in the build container only (needs /root/reference): the reference's unmodified Python
stack (pointnet2_utils / pointnet2_modules / backbone_module / proposal-style vote aggregation)
is imported by oracle/ref_python.py on top of the C oracle `_ext` and evaluated on small seeded
inputs.  The outputs are what `tests/test_golden*.py` compare the oracle port (CPU) and the CUDA
product (GPU) against.  Weights are re-created from `torch.manual_seed(seed)` in the tests; a
checksum stored here guards against RNG drift.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from backtoreality_b200 import scenes  # noqa: E402  (numpy-only generator)
from oracle import ref_python  # noqa: E402


def weight_checksum(module):
    return float(sum(p.detach().double().abs().sum() for p in module.parameters()))


def pattern_like(t):
    """same fixed pseudo-random loss weights as tests/_util.py:pattern_like"""
    i = torch.arange(t.numel(), dtype=torch.float64)
    return torch.sin(i * 12.9898 + 0.5 * torch.cos(i * 0.618)).float().reshape(t.shape)


def sub(t, n=4096):
    """deterministic subsample of a tensor (keeps fixtures small)"""
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step].numpy().copy()


def backbone_case(flavour, C, fp2_out, N, B, seed, train):
    rs = ref_python.RefStack(flavour)
    torch.manual_seed(seed)
    net = rs.backbone_module.Pointnet2Backbone(input_feature_dim=C)
    net.train(train)
    pc = torch.from_numpy(scenes.batch(50, B, N, C=C, kind="room", dup=0.2))
    ep = net(pc)
    out = {"seed": seed, "N": N, "B": B, "C": C, "fp2_out": fp2_out, "train": int(train),
           "wsum": weight_checksum(net),
           "sa1_inds": ep["sa1_inds"].numpy(), "sa2_inds": ep["sa2_inds"].numpy()}
    for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
        out[k] = sub(ep[k])
    # one backward through everything: loss = sum(fp2_features * fixed pattern)
    patt = pattern_like(ep["fp2_features"])
    (ep["fp2_features"] * patt).sum().backward()
    out["g_sa1_l0"] = sub(net.sa1.mlp_module.layer0.conv.weight.grad)
    out["g_sa2_l0"] = sub(net.sa2.mlp_module.layer0.conv.weight.grad)
    out["g_sa4_l2"] = sub(net.sa4.mlp_module.layer2.conv.weight.grad)
    out["g_fp1_l0"] = sub(net.fp1.mlp.layer0.conv.weight.grad)
    out["g_fp2_l1_bn"] = sub(net.fp2.mlp.layer1.bn.bn.weight.grad)
    if train:
        out["rm_sa1_l0"] = net.sa1.mlp_module.layer0.bn.bn.running_mean.numpy().copy()
        out["rv_sa1_l0"] = net.sa1.mlp_module.layer0.bn.bn.running_var.numpy().copy()
    return out


def vote_aggregation_case(seed):
    """ProposalModule.vote_aggregation (proposal_module.py:66-73): xyz requires grad, and the
    `inds=` path of PointnetSAModuleVotes (seed_fps mode, :97-100)."""
    rs = ref_python.RefStack("votenet")
    torch.manual_seed(seed)
    sa = rs.pointnet2_modules.PointnetSAModuleVotes(npoint=64, radius=0.3, nsample=16,
                                                   mlp=[32, 32, 32, 32], use_xyz=True,
                                                   normalize_xyz=True)
    g = torch.Generator().manual_seed(seed + 1)
    xyz = (torch.rand(2, 256, 3, generator=g) * 2.0 + 0.5).requires_grad_(True)
    feats = torch.randn(2, 32, 256, generator=g).requires_grad_(True)
    new_xyz, new_feats, inds = sa(xyz, feats)
    patt = pattern_like(new_feats)
    ((new_feats * patt).sum() + (new_xyz * 0.37).sum()).backward()
    out = {"seed": seed, "wsum": weight_checksum(sa), "inds": inds.numpy(),
           "new_xyz": new_xyz.detach().numpy(), "new_feats": new_feats.detach().numpy(),
           "g_xyz": xyz.grad.numpy().copy(), "g_feats": sub(feats.grad)}
    # explicit inds + features=None (GF3D SA1 style)
    torch.manual_seed(seed)
    sa2 = rs.pointnet2_modules.PointnetSAModuleVotes(npoint=64, radius=0.4, nsample=8,
                                                    mlp=[0, 16, 16], use_xyz=True,
                                                    normalize_xyz=True)
    given = torch.arange(64, dtype=torch.int32).flip(0)[None].repeat(2, 1).contiguous() * 3
    nx, nf, gi = sa2(xyz.detach(), None, given)
    out.update({"wsum2": weight_checksum(sa2), "given": given.numpy(), "nx2": nx.numpy(),
                "nf2": nf.detach().numpy()})
    return out


def fp_case(seed):
    rs = ref_python.RefStack("votenet")
    torch.manual_seed(seed)
    fp = rs.pointnet2_modules.PointnetFPModule(mlp=[48 + 16, 32, 24])
    g = torch.Generator().manual_seed(seed + 1)
    unknown = torch.rand(2, 100, 3, generator=g)
    known = torch.rand(2, 37, 3, generator=g)
    known[:, 5] = known[:, 2]
    uf = torch.randn(2, 16, 100, generator=g).requires_grad_(True)
    kf = torch.randn(2, 48, 37, generator=g).requires_grad_(True)
    y = fp(unknown, known, uf, kf)
    patt = pattern_like(y)
    (y * patt).sum().backward()
    return {"seed": seed, "wsum": weight_checksum(fp), "y": y.detach().numpy(),
            "g_uf": sub(uf.grad), "g_kf": kf.grad.numpy().copy()}


def main():
    if not ref_python.available():
        raise SystemExit("reference tree not found; fixtures can only be generated in the "
                         "build container")
    np.savez_compressed(os.path.join(HERE, "backbone_votenet_eval.npz"),
                        **backbone_case("votenet", 1, 256, 4096, 2, 1234, train=False))
    np.savez_compressed(os.path.join(HERE, "backbone_votenet_train.npz"),
                        **backbone_case("votenet", 1, 256, 4096, 2, 1234, train=True))
    np.savez_compressed(os.path.join(HERE, "backbone_gf3d_train.npz"),
                        **backbone_case("groupfree3d", 0, 288, 3000, 2, 4321, train=True))
    np.savez_compressed(os.path.join(HERE, "vote_aggregation.npz"), **vote_aggregation_case(77))
    np.savez_compressed(os.path.join(HERE, "fp_module.npz"), **fp_case(99))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
