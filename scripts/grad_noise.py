"""How far is a TF32 SharedMLP from an fp32 one, forward and backward?

For each SA-block configuration of the backbone, runs PointnetSAModuleVotes fwd+bwd three ways
on identical inputs/parameters and prints rel-L2 against the fp32 arm:
    fp32     unfused path, cuDNN with TF32 disabled            (the comparison arm)
    cudnn32  unfused path, cuDNN with TF32 ALLOWED             (= what the reference runs)
    fused    the product's tcgen05 TF32 block (csrc/mlp.cu, csrc/mlp_bwd.cu)
and the same for the whole backbone against the golden fixtures.  The numbers set the TF32
tolerances in tests/ (a fused-kernel error comparable to cuDNN's own TF32 error is parity).
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import golden, pattern_like, rel_l2, sub  # noqa: E402
from backtoreality_b200 import fused_sa, scenes  # noqa: E402
from backtoreality_b200.backbone_module import Pointnet2Backbone  # noqa: E402
from backtoreality_b200.pointnet2_modules import PointnetSAModuleVotes  # noqa: E402

dev = torch.device("cuda:0")
CFGS = [dict(N=6000, C=1, npoint=512, radius=0.2, nsample=64, mlp=[1, 64, 64, 128]),
        dict(N=2048, C=128, npoint=1024, radius=0.4, nsample=32, mlp=[128, 128, 128, 256]),
        dict(N=1024, C=256, npoint=256, radius=0.3, nsample=16, mlp=[256, 128, 128, 128]),
        dict(N=3000, C=0, npoint=256, radius=0.3, nsample=16, mlp=[0, 64, 64, 128])]


def set_mode(mode):
    tf32 = mode == "cudnn32"
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
    fused_sa.ENABLED = mode == "fused"


def block(cfg, training):
    torch.manual_seed(13)
    sa = PointnetSAModuleVotes(npoint=cfg["npoint"], radius=cfg["radius"], nsample=cfg["nsample"],
                               mlp=list(cfg["mlp"]), use_xyz=True, normalize_xyz=True).to(dev)
    for blk in sa.mlp_module:
        bn = blk.bn.bn
        bn.weight.data = torch.randn_like(bn.weight) * 0.5 + 0.8
        bn.bias.data = torch.randn_like(bn.bias) * 0.2
        bn.running_mean.data = torch.randn_like(bn.running_mean) * 0.1
        bn.running_var.data = torch.rand_like(bn.running_var) + 0.5
    sa.train(training)
    B = 2
    pc = torch.from_numpy(scenes.batch(71, B, cfg["N"], C=max(cfg["C"], 1), kind="room", dup=0.2)).to(dev)
    res = {}
    for mode in ("fp32", "cudnn32", "fused"):
        set_mode(mode)
        mod = copy.deepcopy(sa)
        xyz = pc[..., :3].contiguous().clone().requires_grad_(True)
        feats = (torch.randn(B, cfg["C"], cfg["N"], device=dev,
                             generator=torch.Generator(device=dev).manual_seed(5)).requires_grad_(True)
                 if cfg["C"] else None)
        new_xyz, y, _ = mod(xyz, feats)
        ((y * pattern_like(y)).sum() + (new_xyz * 0.37).sum()).backward()
        r = {"y": y.detach(), "g_xyz": xyz.grad}
        if feats is not None:
            r["g_feat"] = feats.grad
        for n, p in mod.named_parameters():
            r["g_" + n.replace("mlp_module.", "").replace(".conv.weight", ".W").replace(".bn.bn.weight", ".gamma").replace(".bn.bn.bias", ".beta")] = p.grad
        res[mode] = r
    print("\n== SA block %s training=%s" % (cfg["mlp"], training))
    print("%-22s %12s %12s" % ("tensor", "cudnn-tf32", "fused-tf32"))
    for k in res["fp32"]:
        w = res["fp32"][k].cpu().numpy()
        print("%-22s %12.3e %12.3e" % (k, rel_l2(res["cudnn32"][k].cpu().numpy(), w),
                                       rel_l2(res["fused"][k].cpu().numpy(), w)))


def backbone(fixture):
    g = golden(fixture)
    print("\n== backbone vs golden %s" % fixture)
    print("%-22s %12s %12s %12s" % ("tensor", "fp32", "cudnn-tf32", "fused-tf32"))
    rows = {}
    for mode in ("fp32", "cudnn32", "fused"):
        set_mode(mode)
        torch.manual_seed(int(g["seed"]))
        net = Pointnet2Backbone(input_feature_dim=int(g["C"]), fp2_out=int(g["fp2_out"])).to(dev)
        net.train(bool(g["train"]))
        pc = torch.from_numpy(scenes.batch(50, int(g["B"]), int(g["N"]), C=int(g["C"]), kind="room", dup=0.2)).to(dev)
        ep = net(pc)
        (ep["fp2_features"] * pattern_like(ep["fp2_features"])).sum().backward()
        r = {k: rel_l2(sub(ep[k]), g[k]) for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features")}
        r["g_sa1_l0"] = rel_l2(sub(net.sa1.mlp_module.layer0.conv.weight.grad), g["g_sa1_l0"])
        r["g_sa2_l0"] = rel_l2(sub(net.sa2.mlp_module.layer0.conv.weight.grad), g["g_sa2_l0"])
        r["g_sa4_l2"] = rel_l2(sub(net.sa4.mlp_module.layer2.conv.weight.grad), g["g_sa4_l2"])
        r["g_fp1_l0"] = rel_l2(sub(net.fp1.mlp.layer0.conv.weight.grad), g["g_fp1_l0"])
        r["g_fp2_l1_bn"] = rel_l2(sub(net.fp2.mlp.layer1.bn.bn.weight.grad), g["g_fp2_l1_bn"])
        rows[mode] = r
    for k in rows["fp32"]:
        print("%-22s %12.3e %12.3e %12.3e" % (k, rows["fp32"][k], rows["cudnn32"][k], rows["fused"][k]))


if __name__ == "__main__":
    for cfg in CFGS:
        for tr in (False, True):
            block(cfg, tr)
    for f in ("backbone_votenet_eval.npz", "backbone_votenet_train.npz", "backbone_gf3d_train.npz"):
        backbone(f)
