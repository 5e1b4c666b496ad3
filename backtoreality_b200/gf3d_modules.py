"""GroupFree3D query sampling: the modules between the backbone and the
transformer decoder, same class names / constructor arguments / return tuples / state-dict keys as
/root/reference/detection/GroupFree3D/models/modules.py:16-100, on libb2r:

* `FPSModule` / `GeneralSamplingModule` (modules.py:66-100): FPS (csrc/fps_bucket.cu) + the two
  gathers (xyz rows and the C=288 seed features) -- the reference's three transposing copies
  around `gather_operation` for xyz are one row gather here;
* `PointsObjClsModule` (modules.py:16-43, the KPS objectness head: Conv1d-BN-ReLU x2 + Conv1d) and
  `PositionEmbeddingLearned` (modules.py:46-64: Conv1d-BN-ReLU + Conv1d) on the dense tcgen05 path
  (csrc/dense.cu, K-streamed TF32 GEMM with the BatchNorm statistics in the epilogue); they fall
  back to the reference's own formulation (torch Conv1d / BatchNorm1d on the GPU) for shapes that
  path does not cover (e.g. the 3- or 6-channel input of the position embedding);
* `sample_queries` restates detector.py:156-175 (the 'fps' / 'kps' branches).

The decoder's K/V projections are `nn.Conv1d(288, 288, 1)`; `project_pm` runs one through the same
dense kernel.  The attention itself (transformer.py:36-76, multi_head_attention.py) stays the
reference's torch code: it is outside the set-abstraction path (lowest rank).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dense_mlp, pointnet2_utils


def _dense(x_cm, convs, bns, training):
    """(B,C,n) through [conv-bn-relu]* + conv on the dense path -> (B,C',n), or None."""
    if not x_cm.is_cuda:
        return None
    specs = dense_mlp.layer_specs(convs, bns)
    x_pm = x_cm.transpose(1, 2).contiguous()
    if not dense_mlp.supported(specs, x_pm):
        return None
    B, n, C = x_pm.shape
    out_pm, out_cm = dense_mlp.dense_mlp(x_pm.reshape(B * n, C), specs, training, B, n)
    if out_cm is not None:          # last layer with BatchNorm: channel-major comes with it
        return out_cm
    cout = specs[-1][0].out_channels   # plain last layer: (M, ceil4(C)) raw rows, pad columns undefined
    return out_pm[:, :cout].reshape(B, n, cout).transpose(1, 2).contiguous()


class PointsObjClsModule(nn.Module):
    def __init__(self, seed_feature_dim):
        super().__init__()
        self.in_dim = seed_feature_dim
        self.conv1 = torch.nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.bn1 = torch.nn.BatchNorm1d(self.in_dim)
        self.conv2 = torch.nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.bn2 = torch.nn.BatchNorm1d(self.in_dim)
        self.conv3 = torch.nn.Conv1d(self.in_dim, 1, 1)

    def forward(self, seed_features):
        """seed_features (B,C,num_seed) -> logits (B,1,num_seed)"""
        out = _dense(seed_features, [self.conv1, self.conv2, self.conv3], [self.bn1, self.bn2, None],
                     self.training)
        if out is not None:
            return out
        net = F.relu(self.bn1(self.conv1(seed_features)))
        net = F.relu(self.bn2(self.conv2(net)))
        return self.conv3(net)


class PositionEmbeddingLearned(nn.Module):
    """Absolute pos embedding, learned."""

    def __init__(self, input_channel, num_pos_feats=288):
        super().__init__()
        self.position_embedding_head = nn.Sequential(
            nn.Conv1d(input_channel, num_pos_feats, kernel_size=1),
            nn.BatchNorm1d(num_pos_feats),
            nn.ReLU(inplace=True),
            nn.Conv1d(num_pos_feats, num_pos_feats, kernel_size=1))

    def forward(self, xyz):
        xyz = xyz.transpose(1, 2).contiguous()
        h = self.position_embedding_head
        out = _dense(xyz, [h[0], h[3]], [h[1], None], self.training)
        return out if out is not None else h(xyz)


def _gather_rows(xyz, inds):
    """xyz (B,K,3), inds (B,M) -> (B,M,3); differentiable (reference: transpose + gather_operation +
    transpose, modules.py:80-81)"""
    return torch.gather(xyz, 1, inds.long().unsqueeze(-1).expand(-1, -1, xyz.shape[-1]))


class FPSModule(nn.Module):
    def __init__(self, num_proposal):
        super().__init__()
        self.num_proposal = num_proposal

    def forward(self, xyz, features):
        """xyz (B,K,3), features (B,C,K) -> (new_xyz (B,P,3), new_features (B,C,P), sample_inds)"""
        sample_inds = pointnet2_utils.furthest_point_sample(xyz, self.num_proposal)
        new_xyz = _gather_rows(xyz, sample_inds).contiguous()
        new_features = pointnet2_utils.gather_operation(features, sample_inds).contiguous()
        return new_xyz, new_features, sample_inds


class GeneralSamplingModule(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, xyz, features, sample_inds):
        new_xyz = _gather_rows(xyz, sample_inds).contiguous()
        new_features = pointnet2_utils.gather_operation(features, sample_inds).contiguous()
        return new_xyz, new_features, sample_inds


def sample_queries(end_points, sampling, num_proposal, fps_module=None, points_obj_cls=None,
                   gsample_module=None):
    """detector.py:150-175: seeds = fp2 output; 'fps' samples the queries by FPS, 'kps' by the
    top-k of the learned objectness.  Fills the reference's end_points keys and returns
    (cluster_xyz, cluster_feature)."""
    xyz, features = end_points['fp2_xyz'], end_points['fp2_features']
    end_points['seed_inds'] = end_points['fp2_inds']
    end_points['seed_xyz'] = xyz
    end_points['seed_features'] = features
    if sampling == 'fps':
        xyz, features, sample_inds = fps_module(xyz, features)
    elif sampling == 'kps':
        logits = points_obj_cls(features)
        end_points['seeds_obj_cls_logits'] = logits
        scores = torch.sigmoid(logits).squeeze(1)
        sample_inds = torch.topk(scores, num_proposal)[1].int()
        xyz, features, sample_inds = gsample_module(xyz, features, sample_inds)
    else:
        raise NotImplementedError
    end_points['query_points_xyz'] = xyz
    end_points['query_points_feature'] = features
    end_points['query_points_sample_inds'] = sample_inds
    return xyz, features


def project_pm(conv, x_cm, training=False):
    """decoder_query_proj / decoder_key_proj (detector.py:181-182: nn.Conv1d(C, C, 1)) through the
    dense kernel; falls back to the module itself."""
    out = _dense(x_cm, [conv], [None], training)
    return out if out is not None else conv(x_cm)
