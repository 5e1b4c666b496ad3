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


def centers_case(seed):
    """PointnetSAModuleCenters (pointnet2_modules.py:357-451) as the CenterRefine head uses it
    (backbone_module.py:188-195): external centres, normalize_xyz=False, a ONE-layer MLP."""
    rs = ref_python.RefStack("votenet")
    torch.manual_seed(seed)
    head = rs.pointnet2_modules.PointnetSAModuleCenters(npoint=32, radius=0.8, nsample=16,
                                                        mlp=[64, 32], use_xyz=True,
                                                        normalize_xyz=False)
    g = torch.Generator().manual_seed(seed + 1)
    xyz = (torch.rand(2, 300, 3, generator=g) * 3.0).requires_grad_(True)
    feats = torch.randn(2, 64, 300, generator=g).requires_grad_(True)
    centers = (xyz.detach()[:, :32] + 0.05 * torch.randn(2, 32, 3, generator=g)).requires_grad_(True)
    y = head(xyz, feats, centers)
    patt = pattern_like(y)
    (y * patt).sum().backward()
    out = {"seed": seed, "wsum": weight_checksum(head), "y": y.detach().numpy(),
           "g_xyz": xyz.grad.numpy().copy(), "g_centers": centers.grad.numpy().copy(),
           "g_feats": sub(feats.grad), "g_w": head.mlp_module.layer0.conv.weight.grad.numpy().copy()}
    # GroupFree3D's PointnetSAModuleOffset (pointnet2_modules.py:481-576): three layers
    rg = ref_python.RefStack("groupfree3d")
    torch.manual_seed(seed)
    off = rg.pointnet2_modules.PointnetSAModuleOffset(npoint=32, radius=0.6, nsample=16,
                                                      mlp=[64, 32, 32, 48], use_xyz=True,
                                                      normalize_xyz=True)
    y2 = off(xyz.detach(), feats.detach(), centers.detach())
    out.update({"wsum_off": weight_checksum(off), "y_off": y2.detach().numpy()})
    # ... and its free-function interpolation (:722-730)
    known_xyz = xyz.detach()[:, :40].contiguous()
    out["tni"] = rg.pointnet2_modules.ThreeNNInterpolate(feats.detach()[:, :, :40].contiguous(), known_xyz,
                                                        xyz.detach()[:, 100:180].contiguous()).numpy()
    return out


def jitter_backbone_case(seed):
    """Pointnet2Backbone_jitter (backbone_module.py:136-262) with centres: the reference's
    `.cuda()` on the one-hot classes (:260) is neutralised for the CPU run."""
    rs = ref_python.RefStack("votenet")
    torch.manual_seed(seed)
    net = rs.backbone_module.Pointnet2Backbone_jitter(input_feature_dim=1)
    net.train(True)
    pc = torch.from_numpy(scenes.batch(60, 2, 3000, C=1, kind="room", dup=0.2))
    g = torch.Generator().manual_seed(seed + 1)
    center_xyz = pc[:, :64, :3].clone() + 0.1 * torch.randn(2, 64, 3, generator=g)
    center_cls = torch.randint(0, 22, (2, 64), generator=g)
    with ref_python.cuda_is_identity():
        ep = net(pc, center_xyz, center_cls)
    cf = ep["center_features"]
    patt = pattern_like(cf)
    (cf * patt).sum().backward()
    return {"seed": seed, "wsum": weight_checksum(net), "center_xyz": center_xyz.numpy(),
            "center_cls": center_cls.numpy(), "center_features": cf.detach().numpy(),
            "fp2_features": sub(ep["fp2_features"]),
            "g_ctjt": net.ctjt_head.mlp_module.layer0.conv.weight.grad.numpy().copy(),
            "g_fp2_l1": sub(net.fp2.mlp.layer1.conv.weight.grad),
            "g_sa1_l0": sub(net.sa1.mlp_module.layer0.conv.weight.grad)}


def vote_heads_case(seed, train):
    """VotingModule (voting_module.py:15-65) -> L2 normalisation (votenet.py:93-94) ->
    ProposalModule (proposal_module.py:52-120, seed_fps sampling so that the proposals do not
    depend on the votes' last bits): the callers on either side of the vote aggregation."""
    rs = ref_python.RefStack("votenet")
    torch.manual_seed(seed)
    vgen = rs.voting_module.VotingModule(1, 256)
    msa = np.linspace(0.3, 2.0, 4 * 3).reshape(4, 3).astype(np.float32)
    pnet = rs.proposal_module.ProposalModule(4, 2, 4, msa, 32, "seed_fps")
    for m in list(vgen.modules()) + list(pnet.modules()):
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.momentum = 0.2
            if not train:     # non-trivial running statistics for the eval case
                m.running_mean.data = torch.randn(m.running_mean.shape) * 0.1
                m.running_var.data = torch.rand(m.running_var.shape) + 0.5
    vgen.train(train)
    pnet.train(train)
    g = torch.Generator().manual_seed(seed + 1)
    seed_xyz = (torch.rand(2, 256, 3, generator=g) * 3.0).requires_grad_(True)
    seed_feat = torch.randn(2, 256, 256, generator=g).requires_grad_(True)
    vote_xyz, vote_feat = vgen(seed_xyz, seed_feat)
    vote_feat = vote_feat.div(torch.norm(vote_feat, p=2, dim=1).unsqueeze(1))
    with ref_python.cuda_is_identity():
        ep = pnet(vote_xyz, vote_feat, {"seed_xyz": seed_xyz})
    loss = ((ep["objectness_scores"] * 0.7).sum() + (ep["center"] * 0.3).sum() +
            (ep["size_residuals"] * pattern_like(ep["size_residuals"])).sum() +
            (ep["sem_cls_scores"] * pattern_like(ep["sem_cls_scores"])).sum() +
            (vote_xyz * 0.11).sum())
    loss.backward()
    out = {"seed": seed, "train": int(train), "wsum": weight_checksum(vgen) + weight_checksum(pnet),
           "msa": msa, "vote_xyz": vote_xyz.detach().numpy(), "vote_feat": sub(vote_feat),
           "inds": ep["aggregated_vote_inds"].numpy(),
           "agg_xyz": ep["aggregated_vote_xyz"].detach().numpy(),
           "agg_feat": sub(ep["aggregated_vote_features"]),
           "objectness": ep["objectness_scores"].detach().numpy(),
           "center": ep["center"].detach().numpy(),
           "size_residuals": sub(ep["size_residuals"]), "sem_cls": ep["sem_cls_scores"].detach().numpy(),
           "pred_size": ep["pred_size"].detach().numpy(),
           "g_seed_xyz": seed_xyz.grad.numpy().copy(), "g_seed_feat": sub(seed_feat.grad),
           "g_vgen_c1": sub(vgen.conv1.weight.grad), "g_vgen_c3": sub(vgen.conv3.weight.grad),
           "g_vgen_c3_b": vgen.conv3.bias.grad.numpy().copy(),
           "g_vgen_bn2": vgen.bn2.weight.grad.numpy().copy(),
           "g_pnet_c1": sub(pnet.conv1.weight.grad), "g_pnet_c3_b": pnet.conv3.bias.grad.numpy().copy(),
           "g_pnet_c2_b": pnet.conv2.bias.grad.numpy().copy()}
    if train:
        out["rm_vgen_bn1"] = vgen.bn1.running_mean.numpy().copy()
        out["rv_pnet_bn2"] = pnet.bn2.running_var.numpy().copy()
    return out


def gf3d_query_case(seed):
    """GroupFree3D query sampling (G/models/modules.py:16-100, detector.py:150-175): the KPS
    objectness head, both sampling modules and the learned position embedding, train-mode BN."""
    rg = ref_python.RefStack("groupfree3d")
    m = rg.gf_modules
    torch.manual_seed(seed)
    cls_head = m.PointsObjClsModule(288)
    pos = m.PositionEmbeddingLearned(3, 288)
    pos6 = m.PositionEmbeddingLearned(6, 288)
    fps, gs = m.FPSModule(64), m.GeneralSamplingModule()
    for mod in list(cls_head.modules()) + list(pos.modules()) + list(pos6.modules()):
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.momentum = 0.2
    g = torch.Generator().manual_seed(seed + 1)
    xyz = torch.rand(2, 256, 3, generator=g) * 3.0
    feat = torch.randn(2, 288, 256, generator=g).requires_grad_(True)
    logits = cls_head(feat)
    kps_inds = torch.topk(torch.sigmoid(logits).squeeze(1), 64)[1].int()
    k_xyz, k_feat, _ = gs(xyz, feat, kps_inds)
    f_xyz, f_feat, f_inds = fps(xyz, feat)
    emb = pos(f_xyz)
    emb6 = pos6(torch.cat([f_xyz, f_xyz * 0.5 + 0.1], -1))
    loss = ((logits * pattern_like(logits)).sum() + (k_feat * 0.3).sum() + (f_feat * pattern_like(f_feat)).sum() +
            (emb * pattern_like(emb)).sum() + (emb6 * 0.01).sum())
    loss.backward()
    return {"seed": seed, "wsum": weight_checksum(cls_head) + weight_checksum(pos) + weight_checksum(pos6),
            "logits": logits.detach().numpy(), "kps_inds": kps_inds.numpy(), "k_xyz": k_xyz.detach().numpy(),
            "k_feat": sub(k_feat), "f_inds": f_inds.numpy(), "f_xyz": f_xyz.detach().numpy(),
            "f_feat": sub(f_feat), "emb": sub(emb), "emb6": sub(emb6),
            "g_feat": sub(feat.grad), "g_cls_c1": sub(cls_head.conv1.weight.grad),
            "g_cls_c3_b": cls_head.conv3.bias.grad.numpy().copy(),
            "g_pos_c0": pos.position_embedding_head[0].weight.grad.numpy().copy(),
            "g_pos_c3": sub(pos.position_embedding_head[3].weight.grad),
            "rm_cls_bn1": cls_head.bn1.running_mean.numpy().copy()}


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
    np.savez_compressed(os.path.join(HERE, "centers_offset.npz"), **centers_case(55))
    np.savez_compressed(os.path.join(HERE, "backbone_jitter.npz"), **jitter_backbone_case(2468))
    np.savez_compressed(os.path.join(HERE, "vote_heads_train.npz"), **vote_heads_case(31, True))
    np.savez_compressed(os.path.join(HERE, "vote_heads_eval.npz"), **vote_heads_case(31, False))
    np.savez_compressed(os.path.join(HERE, "gf3d_query_sampling.npz"), **gf3d_query_case(63))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
