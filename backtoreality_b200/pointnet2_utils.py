"""The reference's `pointnet2_utils` autograd API, served by libb2r.so.

Mirror of /root/reference/detection/Votenet/pointnet2/pointnet2_utils.py:51-426 -- same class
names, `.apply` aliases, argument order, return values and non-differentiability markings -- so
code written against the reference (PointnetSAModuleVotes, PointnetFPModule, ProposalModule,
GroupFree3D's sampling modules) runs unchanged:

    furthest_point_sample / FurthestPointSampling   (:51-80)
    gather_operation      / GatherOperation          (:83-117)
    three_nn              / ThreeNN                   (:120-149)
    three_interpolate     / ThreeInterpolate          (:152-206)
    grouping_operation    / GroupingOperation         (:209-257)
    ball_query            / BallQuery                 (:260-291)
    QueryAndGroup, GroupAll                           (:294-426)

Differences, all behind the same call surface:
  * QueryAndGroup runs ONE fused kernel (b2r_query_group_fwd) instead of group(xyz) + in-place
    subtract + in-place divide + group(features) + cat (:347-359): same values, one HBM pass.
    The returned `grouped_xyz` is the first three channels of the fused tensor (a view).
  * backward state is kept with the same ctx attributes as the reference (`for_backwards`,
    `three_interpolate_for_backward`).
  * `sample_uniformly` (:336-345, a CPU double loop never enabled by any caller) is not carried.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _ext


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        """xyz (B,N,3) -> (B,npoint) int32 indices; not differentiable."""
        fps_inds = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(fps_inds)
        return fps_inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        """features (B,C,N), idx (B,npoint) -> (B,C,npoint)."""
        _, C, N = features.size()
        ctx.for_backwards = (idx, C, N)
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        return _ext.gather_points_grad(grad_out.contiguous(), idx, N), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        """unknown (B,n,3), known (B,m,3) -> (dist (B,n,3) L2 distances, idx (B,n,3))."""
        dist2, idx = _ext.three_nn(unknown, known)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        """features (B,c,m), idx (B,n,3), weight (B,n,3) -> (B,c,n)."""
        B, c, m = features.size()
        ctx.three_interpolate_for_backward = (idx, weight, m)
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        grad_features = _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, m)
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        """features (B,C,N), idx (B,npoint,nsample) -> (B,C,npoint,nsample)."""
        _, C, N = features.size()
        ctx.for_backwards = (idx, N)
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, N = ctx.for_backwards
        return _ext.group_points_grad(grad_out.contiguous(), idx, N), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        """radius, nsample, xyz (B,N,3), new_xyz (B,npoint,3) -> (B,npoint,nsample) int32."""
        inds = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class _QueryGroupFused(Function):
    """group(xyz)-centre(/radius) ++ group(features) in one kernel, with its scatter backward."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, features, idx, radius, normalize_xyz):
        C = 0 if features is None else features.size(1)
        ctx.for_backwards = (idx, xyz.size(1), C, radius, normalize_xyz)
        return _ext.query_group(xyz, new_xyz, features, idx, radius, normalize_xyz)

    @staticmethod
    def backward(ctx, grad_out):
        idx, N, C, radius, normalize_xyz = ctx.for_backwards
        need = ctx.needs_input_grad
        gx, gn, gf = _ext.query_group_grad(grad_out.contiguous(), idx, N, C, radius,
                                           normalize_xyz, need[0], need[1], need[2])
        return gx, gn, gf, None, None, None


class QueryAndGroup(nn.Module):
    """Ball query + grouping (reference :294-376).  Returns (B, 3+C, npoint, nsample) with the
    relative (optionally radius-normalised) xyz in the first three channels."""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        if sample_uniformly or ret_unique_cnt:
            raise NotImplementedError(
                "sample_uniformly / ret_unique_cnt: CPU double loop in the reference "
                "(pointnet2_utils.py:336-345), never enabled by any caller; out of scope")
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.sample_uniformly = sample_uniformly
        self.ret_unique_cnt = ret_unique_cnt

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
        if self.use_xyz or features is None:
            new_features = _QueryGroupFused.apply(xyz, new_xyz, features, idx, self.radius,
                                                  self.normalize_xyz)
            grouped_xyz = new_features[:, 0:3]
        else:
            new_features = grouping_operation(features, idx)
            grouped_xyz = None
            if self.ret_grouped_xyz:
                grouped_xyz = _QueryGroupFused.apply(xyz, new_xyz, None, idx, self.radius,
                                                     self.normalize_xyz)
        if self.ret_grouped_xyz:
            return new_features, grouped_xyz
        return new_features


class GroupAll(nn.Module):
    """Groups all features (reference :379-426)."""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            new_features = (torch.cat([grouped_xyz, grouped_features], dim=1)
                            if self.use_xyz else grouped_features)
        else:
            new_features = grouped_xyz
        if self.ret_grouped_xyz:
            return new_features, grouped_xyz
        return new_features
