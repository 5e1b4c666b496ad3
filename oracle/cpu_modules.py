"""CPU restatement of the reference's L1/L2 Python layers on top of the C oracle.

TEST INFRASTRUCTURE ONLY (see b2r_oracle.c).  Used as (i) the checker for module / backbone
outputs and gradients at sizes where no golden fixture is committed, and (ii) the CPU baseline
timed by bench.py (`cpu_baseline`, `--impl reference`): the reference has no CPU path of its own
("CPU not supported", _ext_src/src/sampling.cpp:38-40), so this port is what runs on the host.

Follows, UNFUSED and step by step (paths relative to /root/reference/detection/Votenet/):
  autograd ops         pointnet2/pointnet2_utils.py:51-291
  QueryAndGroup        pointnet2/pointnet2_utils.py:317-376  (group, -=, /=, group, cat)
  SharedMLP            pointnet2/pytorch_utils.py:11-36,67-120
  SA / FP modules      pointnet2/pointnet2_modules.py:210-272, 469-514
  Pointnet2Backbone    models/backbone_module.py:21-133 (GF3D: fp2 width 288)
It is validated against the real reference Python (tests/test_oracle_cpu.py, in the build
container where /root/reference exists) and against the committed golden fixtures.
Parameter names match the reference so state dicts interchange.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import fake_ext as _ext


class _Gather(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.saved = (idx, features.size(2))
        return _ext.gather_points(features.contiguous(), idx)

    @staticmethod
    def backward(ctx, g):
        idx, N = ctx.saved
        return _ext.gather_points_grad(g.contiguous(), idx, N), None


class _Group(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.saved = (idx, features.size(2))
        return _ext.group_points(features.contiguous(), idx)

    @staticmethod
    def backward(ctx, g):
        idx, N = ctx.saved
        return _ext.group_points_grad(g.contiguous(), idx, N), None


class _Interp(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.saved = (idx, weight, features.size(2))
        return _ext.three_interpolate(features.contiguous(), idx, weight.contiguous())

    @staticmethod
    def backward(ctx, g):
        idx, weight, m = ctx.saved
        return _ext.three_interpolate_grad(g.contiguous(), idx, weight.contiguous(), m), None, None


def fps(xyz, npoint):
    return _ext.furthest_point_sampling(xyz.detach().contiguous(), npoint)


def ball_query(radius, nsample, xyz, new_xyz):
    return _ext.ball_query(new_xyz.detach().contiguous(), xyz.detach().contiguous(), radius, nsample)


def three_nn(unknown, known):
    d2, idx = _ext.three_nn(unknown.detach().contiguous(), known.detach().contiguous())
    return torch.sqrt(d2), idx


def query_and_group(xyz, new_xyz, features, radius, nsample, normalize_xyz=True, use_xyz=True):
    """pointnet2_utils.py:317-376, unfused."""
    idx = ball_query(radius, nsample, xyz, new_xyz)
    grouped_xyz = _Group.apply(xyz.transpose(1, 2).contiguous(), idx)
    grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
    if normalize_xyz:
        grouped_xyz = grouped_xyz / radius
    if features is None:
        return grouped_xyz, idx
    grouped_features = _Group.apply(features, idx)
    if use_xyz:
        return torch.cat([grouped_xyz, grouped_features], dim=1), idx
    return grouped_features, idx


def round_tf32(t):
    """`cvt.rna.tf32.f32` (round to nearest, ties away from zero, 10 mantissa bits) of finite fp32
    values: what every forward operand of the product's tensor-core layers goes through
    (csrc/mlp_common.cuh to_tf32)."""
    bits = t.detach().contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1fff).view(torch.float32)


def round_bf16(t):
    """`__float2bfloat16_rn` (round to nearest even): the operands of the fused SA backward."""
    return t.detach().bfloat16().float()


_ROUND = {"fp32": lambda t: t.detach(), "tf32": round_tf32, "bf16": round_bf16}


class _RoundedConv1x1(Function):
    """A 1x1 convolution whose operands are rounded the way the product's kernels round them
    (forward: x and W; backward: grad_out, W and x), products and sums exact (fp64).  With it the
    port is the product's arithmetic up to summation order, so the fused path can be held to a
    tight bound at backbone scale instead of "no worse than cuDNN TF32" (tests/
    test_modules_gpu.py::test_backbone_40k_vs_operand_rounding_emulation)."""

    @staticmethod
    def forward(ctx, x, w, fwd, bwd):
        ctx.save_for_backward(x, w)
        ctx.bwd = bwd
        r = _ROUND[fwd]
        return torch.einsum("oc,bcns->bons", r(w)[:, :, 0, 0].double(), r(x).double()).float()

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        r = _ROUND[ctx.bwd]
        gd = r(g).double()
        dx = torch.einsum("oc,bons->bcns", r(w)[:, :, 0, 0].double(), gd).float()
        dw = torch.einsum("bons,bcns->oc", gd, r(x).double()).float()[:, :, None, None]
        return dx, dw, None, None


def _bc(v):
    return v[None, :, None, None]


class _TopConvBN(Function):
    """The pooled TOP layer of a fused SA block in training mode: 1x1 convolution + BatchNorm as
    one function, because the product's backward does not keep this layer's z -- it recomputes it
    on the tensor cores from the BF16 operands (csrc/mlp_bwd.cu, "top" variant) and uses THAT z in
    the x_hat term of the BatchNorm backward, while the two BatchNorm sums (= the gradients of
    gamma and beta) come from the forward's pooled values.  Everything else as _RoundedConv1x1."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, eps, fwd, bwd):
        r = _ROUND[fwd]
        z = torch.einsum("oc,bcns->bons", r(w)[:, :, 0, 0].double(), r(x).double()).float()
        mean = z.mean((0, 2, 3))
        var = z.var((0, 2, 3), unbiased=False)
        invstd = torch.rsqrt(var + eps)
        ctx.save_for_backward(x, w, gamma, mean, invstd, z)
        ctx.bwd = bwd
        ctx.mark_non_differentiable(mean, var)
        return (z - _bc(mean)) * _bc(invstd * gamma) + _bc(beta), mean, var

    @staticmethod
    def backward(ctx, dy, _gm, _gv):
        x, w, gamma, mean, invstd, z = (t.double() for t in ctx.saved_tensors)
        r = _ROUND[ctx.bwd]
        xb, wb = r(x.float()).double(), r(w.float())[:, :, 0, 0].double()
        n = z.numel() // z.shape[1]
        dy = dy.double()
        s1 = dy.sum((0, 2, 3))
        s2 = (dy * (z - _bc(mean)) * _bc(invstd)).sum((0, 2, 3))
        z_re = torch.einsum("oc,bcns->bons", wb, xb).float().double()   # fp32 accumulators
        xhat = (z_re - _bc(mean)) * _bc(invstd)
        dz = _bc(gamma * invstd) * (dy - _bc(s1) / n - xhat * _bc(s2) / n)
        dz = r(dz.float()).double()
        dx = torch.einsum("oc,bons->bcns", wb, dz).float()
        dw = torch.einsum("bons,bcns->oc", dz, xb).float()[:, :, None, None]
        return dx, dw, s2.float(), s1.float(), None, None, None


class _Conv1x1(nn.Conv2d):
    """nn.Conv2d (same parameters / state-dict keys) with an optional operand-rounding mode:
    `operands = (forward, backward)`, each one of "fp32" | "tf32" | "bf16"; None = plain conv.
    `top_bn` (set by emulate_product_operands on an SA block's last layer) = the BatchNorm2d that
    follows; in training mode conv + BN then run as _TopConvBN and that module passes through."""
    operands = None
    top_bn = None

    def forward(self, x):
        if self.operands is None:
            return super().forward(x)
        bn = self.top_bn
        if bn is not None and bn.training:
            y, mean, var = _TopConvBN.apply(x, self.weight, bn.weight, bn.bias, bn.eps,
                                            *self.operands)
            with torch.no_grad():
                n = x.numel() // x.shape[1]
                m = bn.momentum
                bn.running_mean.mul_(1 - m).add_(m * mean)
                bn.running_var.mul_(1 - m).add_(m * var * n / max(n - 1, 1))
                bn.num_batches_tracked += 1
            return y
        y = _RoundedConv1x1.apply(x, self.weight, *self.operands)
        return y if self.bias is None else y + self.bias.view(1, -1, 1, 1)


def emulate_product_operands(backbone, on=True):
    """Switch a `Backbone` to the operand precisions of the product's kernels:
    fused SA layers TF32 forward / BF16 backward (csrc/mlp.cu, mlp_bwd.cu); a first SA layer with
    <= 8 input channels TF32-rounded operands forward, fp32 backward (csrc/mlp_thin.cu); FP-module
    layers TF32 in both directions (csrc/dense.cu)."""
    for name, m in backbone.named_modules():
        if isinstance(m, SAModuleVotes):
            top = m.mlp_module[len(m.mlp_module) - 1]
            bn = top.bn.bn if hasattr(top, "bn") else None
            top.conv.__dict__["top_bn"] = bn if on else None   # not registered as a submodule
            if bn is not None:
                bn.__dict__.pop("forward", None)
                if on:   # conv + BN run inside _TopConvBN while training
                    bn.forward = (lambda t, bn=bn:
                                  t if bn.training else nn.BatchNorm2d.forward(bn, t))
        if not isinstance(m, _Conv1x1):
            continue
        if not on:
            m.operands = None
        elif ".mlp_module." in "." + name:
            thin = name.endswith("layer0.conv") and m.in_channels <= 8
            m.operands = ("tf32", "fp32") if thin else ("tf32", "bf16")
        else:
            m.operands = ("tf32", "tf32")
    return backbone


def shared_mlp(channels, bn=True):
    """Same nesting/names as pytorch_utils.SharedMLP: layer{i}.conv, layer{i}.bn.bn."""
    seq = nn.Sequential()
    for i in range(len(channels) - 1):
        blk = nn.Sequential()
        conv = _Conv1x1(channels[i], channels[i + 1], kernel_size=(1, 1), bias=not bn)
        nn.init.kaiming_normal_(conv.weight)
        blk.add_module("conv", conv)
        if bn:
            wrap = nn.Sequential()
            wrap.add_module("bn", nn.BatchNorm2d(channels[i + 1]))
            blk.add_module("bn", wrap)
        blk.add_module("activation", nn.ReLU(inplace=True))
        seq.add_module("layer%d" % i, blk)
    return seq


class SAModuleVotes(nn.Module):
    def __init__(self, *, mlp, npoint, radius, nsample, use_xyz=True, normalize_xyz=False, bn=True):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.use_xyz, self.normalize_xyz = use_xyz, normalize_xyz
        spec = list(mlp)
        if use_xyz:
            spec[0] += 3
        self.mlp_module = shared_mlp(spec, bn=bn)

    def forward(self, xyz, features=None, inds=None):
        if inds is None:
            inds = fps(xyz, self.npoint)
        new_xyz = _Gather.apply(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
        grouped, idx = query_and_group(xyz, new_xyz, features, self.radius, self.nsample,
                                       self.normalize_xyz, self.use_xyz)
        y = self.mlp_module(grouped)
        y = F.max_pool2d(y, kernel_size=[1, y.size(3)]).squeeze(-1)
        return new_xyz, y, inds


class FPModule(nn.Module):
    def __init__(self, *, mlp, bn=True):
        super().__init__()
        self.mlp = shared_mlp(list(mlp), bn=bn)

    def forward(self, unknown, known, unknow_feats, known_feats):
        dist, idx = three_nn(unknown, known)
        dist_recip = 1.0 / (dist + 1e-8)
        weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
        interp = _Interp.apply(known_feats, idx, weight)
        x = torch.cat([interp, unknow_feats], dim=1) if unknow_feats is not None else interp
        return self.mlp(x.unsqueeze(-1)).squeeze(-1)


class Backbone(nn.Module):
    def __init__(self, input_feature_dim=0, fp2_out=256):
        super().__init__()
        self.sa1 = SAModuleVotes(npoint=2048, radius=0.2, nsample=64,
                                 mlp=[input_feature_dim, 64, 64, 128], normalize_xyz=True)
        self.sa2 = SAModuleVotes(npoint=1024, radius=0.4, nsample=32,
                                 mlp=[128, 128, 128, 256], normalize_xyz=True)
        self.sa3 = SAModuleVotes(npoint=512, radius=0.8, nsample=16,
                                 mlp=[256, 128, 128, 256], normalize_xyz=True)
        self.sa4 = SAModuleVotes(npoint=256, radius=1.2, nsample=16,
                                 mlp=[256, 128, 128, 256], normalize_xyz=True)
        self.fp1 = FPModule(mlp=[512, 256, 256])
        self.fp2 = FPModule(mlp=[512, 256, fp2_out])

    def forward(self, pc):
        xyz = pc[..., 0:3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        ep = {}
        xyz, features, inds = self.sa1(xyz, features)
        ep['sa1_inds'], ep['sa1_xyz'], ep['sa1_features'] = inds, xyz, features
        xyz, features, inds = self.sa2(xyz, features)
        ep['sa2_inds'], ep['sa2_xyz'], ep['sa2_features'] = inds, xyz, features
        xyz, features, inds = self.sa3(xyz, features)
        ep['sa3_xyz'], ep['sa3_features'] = xyz, features
        xyz, features, inds = self.sa4(xyz, features)
        ep['sa4_xyz'], ep['sa4_features'] = xyz, features
        features = self.fp1(ep['sa3_xyz'], ep['sa4_xyz'], ep['sa3_features'], ep['sa4_features'])
        features = self.fp2(ep['sa2_xyz'], ep['sa3_xyz'], ep['sa2_features'], features)
        ep['fp2_features'] = features
        ep['fp2_xyz'] = ep['sa2_xyz']
        ep['fp2_inds'] = ep['sa1_inds'][:, 0:ep['fp2_xyz'].shape[1]]
        return ep


class VoteNetCPU(nn.Module):
    """CPU port of BASELINE.json config 2's model: backbone -> VotingModule -> L2 normalise ->
    vote aggregation + proposal head (models/votenet.py:67-100, voting_module.py:38-65,
    proposal_module.py:84-119).  Same attribute names as the reference, so state dicts
    interchange with backtoreality_b200.votenet.VoteNet."""

    def __init__(self, num_class=22, num_heading_bin=1, num_size_cluster=22, input_feature_dim=1,
                 num_proposal=256):
        super().__init__()
        self.backbone_net = Backbone(input_feature_dim=input_feature_dim)
        self.vgen = nn.Module()
        self.vgen.conv1 = nn.Conv1d(256, 256, 1)
        self.vgen.conv2 = nn.Conv1d(256, 256, 1)
        self.vgen.conv3 = nn.Conv1d(256, 3 + 256, 1)
        self.vgen.bn1 = nn.BatchNorm1d(256)
        self.vgen.bn2 = nn.BatchNorm1d(256)
        self.pnet = nn.Module()
        self.pnet.vote_aggregation = SAModuleVotes(npoint=num_proposal, radius=0.3, nsample=16,
                                                   mlp=[256, 128, 128, 128], normalize_xyz=True)
        self.pnet.conv1 = nn.Conv1d(128, 128, 1)
        self.pnet.conv2 = nn.Conv1d(128, 128, 1)
        self.pnet.conv3 = nn.Conv1d(
            128, 2 + 3 + num_heading_bin * 2 + num_size_cluster * 4 + num_class, 1)
        self.pnet.bn1 = nn.BatchNorm1d(128)
        self.pnet.bn2 = nn.BatchNorm1d(128)

    def forward(self, inputs):
        ep = self.backbone_net(inputs['point_clouds'])
        seed_xyz, seed_features = ep['fp2_xyz'], ep['fp2_features']
        v = self.vgen
        net = F.relu(v.bn1(v.conv1(seed_features)))
        net = F.relu(v.bn2(v.conv2(net)))
        net = v.conv3(net).transpose(2, 1)
        vote_xyz = (seed_xyz + net[:, :, 0:3]).contiguous()
        vote_features = (seed_features.transpose(2, 1) + net[:, :, 3:]).transpose(2, 1).contiguous()
        vote_features = vote_features.div(torch.norm(vote_features, p=2, dim=1).unsqueeze(1))
        ep['vote_xyz'], ep['vote_features'] = vote_xyz, vote_features
        p = self.pnet
        xyz, features, inds = p.vote_aggregation(vote_xyz, vote_features)
        ep['aggregated_vote_xyz'], ep['aggregated_vote_features'] = xyz, features
        ep['aggregated_vote_inds'] = inds
        net = F.relu(p.bn1(p.conv1(features)))
        net = F.relu(p.bn2(p.conv2(net)))
        ep['proposal_scores_raw'] = p.conv3(net)
        return ep


def nn_distance(pc1, pc2, l1smooth=False, delta=1.0, l1=False):
    """The reference's tile-and-min formulation (utils/nn_distance.py:34-61), restated: the checker
    of backtoreality_b200.nn_distance."""
    N, M = pc1.shape[1], pc2.shape[1]
    diff = pc1.unsqueeze(2).repeat(1, 1, M, 1) - pc2.unsqueeze(1).repeat(1, N, 1, 1)
    if l1smooth:
        a = torch.abs(diff)
        q = torch.clamp(a, max=delta)
        cost = torch.sum(0.5 * q ** 2 + delta * (a - q), dim=-1)
    elif l1:
        cost = torch.sum(torch.abs(diff), dim=-1)
    else:
        cost = torch.sum(diff ** 2, dim=-1)
    dist1, idx1 = torch.min(cost, dim=2)
    dist2, idx2 = torch.min(cost, dim=1)
    return dist1, idx1, dist2, idx2
