"""PointNet++ set-abstraction / feature-propagation modules on the B200-native ops.

Mirrors the call surface of /root/reference/detection/Votenet/pointnet2/pointnet2_modules.py that
the detectors instantiate:

    PointnetSAModuleVotes    (:164-272)   used by Pointnet2Backbone and ProposalModule
    PointnetFPModule         (:454-514)   used by Pointnet2Backbone
    PointnetSAModuleCenters  (:357-451)   used by the CenterRefine backbones
    PointnetSAModuleOffset   GroupFree3D/pointnet2/pointnet2_modules.py:481-576 (imported by
                             models/detector_DA.py:16)
    ThreeNNInterpolate       GroupFree3D/pointnet2/pointnet2_modules.py:722-730

Constructor keyword arguments, forward signatures, return tuples, sub-module names
(`grouper`, `mlp_module`, `mlp`) and therefore state-dict keys are the reference's.  The
reference's quirk of mutating the caller's `mlp` list (`mlp_spec[0] += 3`, :204-206) is kept.
MSG / LFP / Rlt variants are never instantiated by any training script and are not carried
(SURVEY.md 2a row 3).
"""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dense_mlp, fused_sa
from . import pointnet2_utils
from . import pytorch_utils as pt_utils


def _pool(new_features, grouped_xyz, pooling, sigma, nsample):
    """(B,C,npoint,nsample) -> (B,C,npoint): the reference's three pooling modes (:254-267)."""
    if pooling == 'max':
        new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
    elif pooling == 'avg':
        new_features = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)])
    elif pooling == 'rbf':
        rbf = torch.exp(-1 * grouped_xyz.pow(2).sum(1, keepdim=False) / (sigma ** 2) / 2)
        new_features = torch.sum(new_features * rbf.unsqueeze(1), -1, keepdim=True) / float(nsample)
    return new_features.squeeze(-1)


class _SAVotesBase(nn.Module):
    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None,
                 nsample: int = None, bn: bool = True, use_xyz: bool = True,
                 pooling: str = 'max', sigma: float = None, normalize_xyz: bool = False,
                 sample_uniformly: bool = False, ret_unique_cnt: bool = False):
        super().__init__()
        self.npoint = npoint
        self.radius = radius
        self.nsample = nsample
        self.pooling = pooling
        self.mlp_module = None
        self.use_xyz = use_xyz
        self.sigma = sigma
        if self.sigma is None:
            self.sigma = self.radius / 2
        self.normalize_xyz = normalize_xyz
        self.ret_unique_cnt = ret_unique_cnt
        # cap of the fused kernels' persistent grids when no `geometry` carries one: 0 = every
        # SM, int or (forward, backward) otherwise (see fused_sa.sa_block)
        self.sm_limit = 0

        if npoint is not None:
            self.grouper = pointnet2_utils.QueryAndGroup(
                radius, nsample, use_xyz=use_xyz, ret_grouped_xyz=True,
                normalize_xyz=normalize_xyz, sample_uniformly=sample_uniformly,
                ret_unique_cnt=ret_unique_cnt)
        else:
            self.grouper = pointnet2_utils.GroupAll(use_xyz, ret_grouped_xyz=True)

        mlp_spec = mlp
        if use_xyz and len(mlp_spec) > 0:
            mlp_spec[0] += 3  # mutates the caller's list, exactly like the reference
        self.mlp_module = pt_utils.SharedMLP(mlp_spec, bn=bn)

    def fusable(self, xyz):
        g = self.grouper
        return (fused_sa.ENABLED and xyz.is_cuda and self.pooling == 'max'
                and isinstance(g, pointnet2_utils.QueryAndGroup) and g.use_xyz
                and not g.sample_uniformly and not g.ret_unique_cnt)

    def _abstract(self, xyz, new_xyz, features, idx=None, sm_limit=0, features_pm=None,
                  want_pm=False, plan=None):
        """-> (new_features (B,C,npoint), the same point-major or None).  features_pm / want_pm:
        point-major hand-over between consecutive fused blocks (fused_sa.sa_block)."""
        if self.fusable(xyz):
            # fused path: ball query, then ONE tcgen05 block for group -> relative xyz -> MLP ->
            # BN/ReLU -> max-pool (csrc/mlp.cu, csrc/mlp_bwd.cu); no (B,C,npoint,nsample) tensor
            if idx is None:
                idx = pointnet2_utils.ball_query(self.radius, self.nsample, xyz, new_xyz)
            if fused_sa.supported(self.mlp_module, xyz, features, idx, self.pooling):
                return fused_sa.sa_block(xyz, new_xyz, features, idx, self.radius,
                                         self.normalize_xyz, self.mlp_module, self.training,
                                         sm_limit=sm_limit, features_pm=features_pm,
                                         want_pm=want_pm, plan=plan)
        grouped_features, grouped_xyz = self.grouper(xyz, new_xyz, features)
        new_features = self.mlp_module(grouped_features)  # (B, mlp[-1], npoint, nsample)
        return _pool(new_features, grouped_xyz, self.pooling, self.sigma, self.nsample), None


class PointnetSAModuleVotes(_SAVotesBase):
    """Set abstraction that also returns the sampled indices (reference :164-272).

    forward(xyz (B,N,3), features (B,C,N) | None, inds (B,npoint) int32 | None)
        -> new_xyz (B,npoint,3), new_features (B,mlp[-1],npoint), inds (B,npoint)
    """

    def forward_pm(self, xyz, features=None, inds=None, features_pm=None):
        """`forward` for callers that hand features over POINT-major: features_pm (B,N,C) is used
        instead of `features` by the fused block, and the block's output comes back in both
        layouts: -> (new_xyz, new_features (B,C,npoint), inds, new_features_pm (B,npoint,C) | None)
        (None when the unfused path ran)."""
        xyz_flipped = xyz.transpose(1, 2).contiguous()
        if inds is None:
            inds = pointnet2_utils.furthest_point_sample(xyz, self.npoint)
        else:
            assert (inds.shape[1] == self.npoint)
        new_xyz = pointnet2_utils.gather_operation(xyz_flipped, inds).transpose(1, 2).contiguous()
        new_features, out_pm = self._abstract(xyz, new_xyz, features, sm_limit=self.sm_limit,
                                              features_pm=features_pm, want_pm=True)
        return new_xyz, new_features, inds, out_pm

    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None,
                inds: torch.Tensor = None, geometry: dict = None):
        if geometry is not None:
            # FPS / centre gather / ball query were done ahead of time on a geometry stream
            # (backbone_module.Pointnet2Backbone.geometry_prepass): wait for them, run the MLP
            if geometry.get("event") is not None:
                torch.cuda.current_stream().wait_event(geometry["event"])
            plan = None
            if fused_sa.COMPACT and "cmeta" in geometry:   # pad-free position space, pre-built
                plan = {k: geometry[k] for k in fused_sa.PLAN_KEYS}
            new_features, out_pm = self._abstract(
                xyz, geometry["new_xyz"], features, idx=geometry["idx"],
                sm_limit=geometry.get("sm_limit", 0), features_pm=geometry.get("features_pm"),
                want_pm=geometry.get("want_pm", False), plan=plan)
            if out_pm is not None:
                geometry["out_pm"] = out_pm      # for the next block (Pointnet2Backbone.forward)
            return geometry["new_xyz"], new_features, geometry["inds"]
        xyz_flipped = xyz.transpose(1, 2).contiguous()
        if inds is None:
            inds = pointnet2_utils.furthest_point_sample(xyz, self.npoint)
        else:
            assert (inds.shape[1] == self.npoint)
        new_xyz = pointnet2_utils.gather_operation(
            xyz_flipped, inds
        ).transpose(1, 2).contiguous() if self.npoint is not None else None
        new_features, _ = self._abstract(xyz, new_xyz, features, sm_limit=self.sm_limit)
        return new_xyz, new_features, inds


class PointnetSAModuleCenters(_SAVotesBase):
    """Set abstraction around externally supplied centres (reference :357-451).

    forward(xyz (B,N,3), features (B,C,N), centers (B,npoint,3)) -> new_features (B,mlp[-1],npoint)
    """

    def forward(self, xyz: torch.Tensor, features: torch.Tensor, centers: torch.Tensor):
        return self._abstract(xyz, centers, features)[0]


class PointnetSAModuleOffset(_SAVotesBase):
    """Set abstraction around externally supplied points (GroupFree3D reference
    pointnet2/pointnet2_modules.py:481-576): same computation as PointnetSAModuleCenters with the
    third argument named `new_xyz`, plus the `ret_unique_cnt` return form (new_features,
    unique_cnt), which -- like in PointnetSAModuleVotes -- runs the unfused QueryAndGroup.

    forward(xyz (B,N,3), features (B,C,N), new_xyz (B,npoint,3)) -> new_features (B,mlp[-1],npoint)
    """

    def forward(self, xyz: torch.Tensor, features: torch.Tensor, new_xyz: torch.Tensor):
        if self.ret_unique_cnt:
            grouped_features, grouped_xyz, unique_cnt = self.grouper(xyz, new_xyz, features)
            new_features = self.mlp_module(grouped_features)
            return _pool(new_features, grouped_xyz, self.pooling, self.sigma, self.nsample), unique_cnt
        return self._abstract(xyz, new_xyz, features)[0]


def ThreeNNInterpolate(known_feats, known_xyz, unknown_xyz):
    """Inverse-distance 3-NN interpolation of known_feats (B,C,m) at unknown_xyz (B,n,3)
    (GroupFree3D reference pointnet2/pointnet2_modules.py:722-730; the front half of
    PointnetFPModule.forward as a free function) -> (B,C,n)."""
    idx, weight = PointnetFPModule.interpolation_weights(unknown_xyz, known_xyz)
    return pointnet2_utils.three_interpolate(known_feats, idx, weight)


class PointnetFPModule(nn.Module):
    """Feature propagation: inverse-distance 3-NN interpolation + SharedMLP (reference :454-514).

    forward(unknown (B,n,3), known (B,m,3), unknow_feats (B,C1,n), known_feats (B,C2,m))
        -> (B, mlp[-1], n)
    """

    def __init__(self, *, mlp: List[int], bn: bool = True):
        super().__init__()
        self.mlp = pt_utils.SharedMLP(mlp, bn=bn)

    @staticmethod
    def interpolation_weights(unknown, known):
        """three_nn + the reference's inverse-distance weights (:486-490): (idx, weight), both
        (B,n,3).  They depend on coordinates only, so a geometry pre-pass may compute them ahead
        of time and hand them to `forward(..., interp=)`."""
        dist, idx = pointnet2_utils.three_nn(unknown, known)
        dist_recip = 1.0 / (dist + 1e-8)
        norm = torch.sum(dist_recip, dim=2, keepdim=True)
        return idx, dist_recip / norm

    def _specs(self):
        return dense_mlp.layer_specs([blk.conv for blk in self.mlp], [blk.bn.bn if hasattr(blk, "bn")
                                                                    else None for blk in self.mlp])

    def forward_pm(self, unknown, known, unknow_feats_pm, known_feats_pm, interp=None):
        """The module on POINT-major features: unknow_feats_pm (B,n,C1) | None, known_feats_pm
        (B,m,C2) -> (new_features (B,mlp[-1],n), the same point-major (B,n,mlp[-1])), or None when
        the dense tcgen05 path does not cover this module (callers then use `forward`).
        Interpolation + concatenation run as one pass, the SharedMLP on csrc/dense.cu."""
        specs = self._specs()
        if known is None or not dense_mlp.supported(specs, known_feats_pm) or \
                any(bn is None for _, bn in specs):
            return None
        idx, weight = interp if interp is not None else self.interpolation_weights(unknown, known)
        x0 = dense_mlp.interp_cat(known_feats_pm, unknow_feats_pm, idx, weight)
        B, n = idx.shape[0], idx.shape[1]
        out_pm, out_cm = dense_mlp.dense_mlp(x0, specs, self.training, B, n)
        return out_cm, out_pm

    def forward(self, unknown, known, unknow_feats, known_feats, interp=None):
        if known is not None and known_feats.is_cuda and dense_mlp.enabled():
            # fused path: same maths, point-major inside (two transposing copies at the boundary)
            res = self.forward_pm(unknown, known,
                                  unknow_feats.transpose(1, 2).contiguous()
                                  if unknow_feats is not None else None,
                                  known_feats.transpose(1, 2).contiguous(), interp=interp)
            if res is not None:
                return res[0]
        if known is not None:
            idx, weight = interp if interp is not None else self.interpolation_weights(unknown, known)
            interpolated_feats = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated_feats = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))

        if unknow_feats is not None:
            new_features = torch.cat([interpolated_feats, unknow_feats], dim=1)
        else:
            new_features = interpolated_feats

        new_features = self.mlp(new_features.unsqueeze(-1))
        return new_features.squeeze(-1)
