"""VoteNet's callers of the hot path: VotingModule, ProposalModule (vote aggregation), VoteNet.

These sit on either side of the set-abstraction path and are plain torch in
the reference; they are restated here only so that BASELINE.json's config 2 ("backbone + vote
head fwd/bwd") can run end to end on the B200-native ops with the reference's attribute names
(`backbone_net`, `vgen`, `pnet`, `vote_aggregation`, conv1..3, bn1..2) and hence its checkpoint
keys.  Mirrors /root/reference/detection/Votenet/models/voting_module.py:15-65,
models/proposal_module.py:18-120 and models/votenet.py:28-100.  `decode_scores` is
device-agnostic (the reference hard-codes `.cuda()`, proposal_module.py:40).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dense_mlp, fused_sa, pointnet2_utils
from .backbone_module import Pointnet2Backbone
from .pointnet2_modules import PointnetSAModuleVotes


class VotingModule(nn.Module):
    """Seeds -> votes: three 1x1 Conv1d, xyz offset + residual features (voting_module.py:15-65)."""

    def __init__(self, vote_factor, seed_feature_dim):
        super().__init__()
        self.vote_factor = vote_factor
        self.in_dim = seed_feature_dim
        self.out_dim = self.in_dim
        self.conv1 = nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.conv2 = nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.conv3 = nn.Conv1d(self.in_dim, (3 + self.out_dim) * self.vote_factor, 1)
        self.bn1 = nn.BatchNorm1d(self.in_dim)
        self.bn2 = nn.BatchNorm1d(self.in_dim)

    def forward_normalized(self, seed_xyz, seed_features, seed_features_pm=None):
        """forward + VoteNet's L2 feature normalisation (votenet.py:93-94) on the dense tcgen05
        path (csrc/dense.cu, csrc/heads.cu): -> (vote_xyz (B,n,3), vote_features (B,C,n) normalised,
        the same point-major (B,n,C)), or None when that path does not cover this module.
        seed_features_pm (B,n,C): the seed features point-major, when the caller has them."""
        if self.vote_factor != 1 or not seed_xyz.is_cuda:
            return None
        specs = dense_mlp.layer_specs([self.conv1, self.conv2, self.conv3], [self.bn1, self.bn2, None])
        if seed_features_pm is None:
            seed_features_pm = seed_features.transpose(1, 2).contiguous()
        if not dense_mlp.supported(specs, seed_features_pm):
            return None
        B, n, C = seed_features_pm.shape
        net, _ = dense_mlp.dense_mlp(seed_features_pm.reshape(B * n, C), specs, self.training, B, n)
        vote_xyz, vf_pm = dense_mlp.vote_tail(net, seed_xyz, seed_features_pm)
        return vote_xyz, vf_pm.transpose(1, 2).contiguous(), vf_pm

    def forward(self, seed_xyz, seed_features):
        B, num_seed = seed_xyz.shape[0], seed_xyz.shape[1]
        num_vote = num_seed * self.vote_factor
        net = F.relu(self.bn1(self.conv1(seed_features)))
        net = F.relu(self.bn2(self.conv2(net)))
        net = self.conv3(net)
        net = net.transpose(2, 1).view(B, num_seed, self.vote_factor, 3 + self.out_dim)
        vote_xyz = (seed_xyz.unsqueeze(2) + net[:, :, :, 0:3]).contiguous().view(B, num_vote, 3)
        vote_features = seed_features.transpose(2, 1).unsqueeze(2) + net[:, :, :, 3:]
        vote_features = vote_features.contiguous().view(B, num_vote, self.out_dim)
        return vote_xyz, vote_features.transpose(2, 1).contiguous()


def decode_scores(net, end_points, num_class, num_heading_bin, num_size_cluster, mean_size_arr):
    """Slice the proposal head output into named predictions (proposal_module.py:18-50)."""
    nt = net.transpose(2, 1)
    B, P = nt.shape[0], nt.shape[1]
    NH, NS = num_heading_bin, num_size_cluster
    end_points['objectness_scores'] = nt[:, :, 0:2]
    end_points['center'] = end_points['aggregated_vote_xyz'] + nt[:, :, 2:5]
    end_points['heading_scores'] = nt[:, :, 5:5 + NH]
    hrn = nt[:, :, 5 + NH:5 + NH * 2]
    end_points['heading_residuals_normalized'] = hrn
    end_points['heading_residuals'] = hrn * (np.pi / NH)
    size_scores = nt[:, :, 5 + NH * 2:5 + NH * 2 + NS]
    srn = nt[:, :, 5 + NH * 2 + NS:5 + NH * 2 + NS * 4].view([B, P, NS, 3])
    end_points['size_scores'] = size_scores
    end_points['size_residuals_normalized'] = srn
    if isinstance(mean_size_arr, torch.Tensor):   # device-resident copy kept by ProposalModule
        msa = mean_size_arr[None, None]
    else:
        msa = torch.from_numpy(np.asarray(mean_size_arr, np.float32)).to(net.device)[None, None]
    end_points['size_residuals'] = srn * msa
    size_recover = msa + end_points['size_residuals']
    cls = torch.argmax(size_scores, -1).unsqueeze(-1).unsqueeze(-1).repeat(1, 1, 1, 3)
    end_points['pred_size'] = torch.gather(size_recover, 2, cls).squeeze(2)
    end_points['sem_cls_scores'] = nt[:, :, 5 + NH * 2 + NS * 4:]
    return end_points


class ProposalModule(nn.Module):
    """Vote aggregation (a PointnetSAModuleVotes) + 3-conv proposal head
    (proposal_module.py:52-120)."""

    def __init__(self, num_class, num_heading_bin, num_size_cluster, mean_size_arr, num_proposal,
                 sampling, seed_feat_dim=256):
        super().__init__()
        self.num_class = num_class
        self.num_heading_bin = num_heading_bin
        self.num_size_cluster = num_size_cluster
        self.mean_size_arr = mean_size_arr
        self.num_proposal = num_proposal
        self.sampling = sampling
        self.seed_feat_dim = seed_feat_dim
        self.vote_aggregation = PointnetSAModuleVotes(
            npoint=self.num_proposal, radius=0.3, nsample=16,
            mlp=[self.seed_feat_dim, 128, 128, 128], use_xyz=True, normalize_xyz=True)
        self.conv1 = nn.Conv1d(128, 128, 1)
        self.conv2 = nn.Conv1d(128, 128, 1)
        self.conv3 = nn.Conv1d(
            128, 2 + 3 + num_heading_bin * 2 + num_size_cluster * 4 + self.num_class, 1)
        self.bn1 = nn.BatchNorm1d(128)
        self.bn2 = nn.BatchNorm1d(128)
        # device-resident mean sizes (the reference re-uploads the numpy array every forward,
        # proposal_module.py:40: a host->device copy per step that also forbids graph capture);
        # non-persistent, so the state dict stays the reference's
        self.register_buffer("_mean_size", torch.from_numpy(
            np.asarray(mean_size_arr, np.float32)).clone(), persistent=False)

    def forward(self, xyz, features, end_points, features_pm=None):
        """features_pm (not in the reference's signature): the vote features point-major
        (B,num_vote,C), when the caller has them (VotingModule.forward_normalized)."""
        if self.sampling == 'vote_fps':
            sample_inds = None
        elif self.sampling == 'seed_fps':
            sample_inds = pointnet2_utils.furthest_point_sample(end_points['seed_xyz'],
                                                                self.num_proposal)
        elif self.sampling == 'random':
            num_seed = end_points['seed_xyz'].shape[1]
            B = end_points['seed_xyz'].shape[0]
            sample_inds = torch.randint(0, num_seed, (B, self.num_proposal), dtype=torch.int,
                                        device=xyz.device)
        else:
            raise ValueError('Unknown sampling strategy: %s' % (self.sampling,))
        xyz, features, fps_inds, agg_pm = self.vote_aggregation.forward_pm(xyz, features, sample_inds,
                                                                          features_pm)
        if sample_inds is None:
            sample_inds = fps_inds
        end_points['aggregated_vote_xyz'] = xyz
        end_points['aggregated_vote_features'] = features
        end_points['aggregated_vote_inds'] = sample_inds

        net = None
        if features.is_cuda and dense_mlp.enabled():
            # proposal head on the dense tcgen05 path, point-major (csrc/dense.cu)
            specs = dense_mlp.layer_specs([self.conv1, self.conv2, self.conv3],
                                          [self.bn1, self.bn2, None])
            if agg_pm is None:
                agg_pm = features.transpose(1, 2).contiguous()
            if dense_mlp.supported(specs, agg_pm):
                B, P, C = agg_pm.shape
                net_pm, _ = dense_mlp.dense_mlp(agg_pm.reshape(B * P, C), specs, self.training, B, P)
                Cout = self.conv3.out_channels
                net = net_pm[:, :Cout].reshape(B, P, Cout).transpose(1, 2)    # (B, Cout, P) view
        if net is None:
            net = F.relu(self.bn1(self.conv1(features)))
            net = F.relu(self.bn2(self.conv2(net)))
            net = self.conv3(net)
        end_points['proposal_scores_raw'] = net
        return decode_scores(net, end_points, self.num_class, self.num_heading_bin,
                             self.num_size_cluster, self._mean_size)


class VoteNet(nn.Module):
    """backbone_net -> vgen -> L2-normalised vote features -> pnet (votenet.py:28-100)."""

    def __init__(self, num_class, num_heading_bin, num_size_cluster, mean_size_arr,
                 input_feature_dim=0, num_proposal=128, vote_factor=1, sampling='vote_fps'):
        super().__init__()
        self.num_class = num_class
        self.num_heading_bin = num_heading_bin
        self.num_size_cluster = num_size_cluster
        self.mean_size_arr = mean_size_arr
        assert (mean_size_arr.shape[0] == self.num_size_cluster)
        self.input_feature_dim = input_feature_dim
        self.num_proposal = num_proposal
        self.vote_factor = vote_factor
        self.sampling = sampling
        self.backbone_net = Pointnet2Backbone(input_feature_dim=self.input_feature_dim)
        self.vgen = VotingModule(self.vote_factor, 256)
        self.pnet = ProposalModule(num_class, num_heading_bin, num_size_cluster, mean_size_arr,
                                   num_proposal, sampling)

    def forward(self, inputs):
        pc = inputs['point_clouds']
        prepacked = pc.is_cuda and fused_sa.ENABLED
        if prepacked:
            # operand images of all five fused blocks, packed on a side stream (fused_sa.prepack)
            bb = self.backbone_net
            fused_sa.prepack([m.mlp_module for m in (bb.sa1, bb.sa2, bb.sa3, bb.sa4,
                                                     self.pnet.vote_aggregation)])
            # ... and of the dense FP / voting / proposal layers (dense_mlp.prepack)
            dense_mlp.prepack([blk.conv for blk in list(bb.fp1.mlp) + list(bb.fp2.mlp)] +
                              [self.vgen.conv1, self.vgen.conv2, self.vgen.conv3,
                               self.pnet.conv1, self.pnet.conv2, self.pnet.conv3])
        # 'geometry' (optional, not in the reference): sa1..sa4 indices computed ahead of time
        end_points = self.backbone_net(pc, {'_prepacked': True} if prepacked else {},
                                       geometry=inputs.get('geometry'))
        xyz = end_points['fp2_xyz']
        features = end_points['fp2_features']
        end_points['seed_inds'] = end_points['fp2_inds']
        end_points['seed_xyz'] = xyz
        end_points['seed_features'] = features
        res = None
        if dense_mlp.enabled() and xyz.is_cuda:
            # voting MLP + offset / residual split + normalisation on the dense tcgen05 path
            res = self.vgen.forward_normalized(xyz, features, end_points.get('fp2_features_pm'))
        if res is not None:
            xyz, features, features_pm = res
        else:
            xyz, features = self.vgen(xyz, features)
            features_norm = torch.norm(features, p=2, dim=1)
            features = features.div(features_norm.unsqueeze(1))
            features_pm = None
        end_points['vote_xyz'] = xyz
        end_points['vote_features'] = features
        end_points = self.pnet(xyz, features, end_points, features_pm=features_pm)
        if prepacked:
            fused_sa.prepack_join()
            dense_mlp.prepack_join()
        return end_points


class _GradReverse(torch.autograd.Function):
    """Identity in forward, -1 x gradient in backward (votenet_DA.py:30-44)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output * -1.0


def grad_reverse(x):
    return _GradReverse.apply(x)


class VoteNet_DA(VoteNet):
    """VoteNet with the global / local domain discriminators of the "Back to Reality" training
    step (reference models/votenet_DA.py:47-176; BASELINE.json configs[2]): same backbone, voting
    and proposal modules -- the hot path -- plus two small gradient-reversed heads on the seed
    features and on the aggregated vote features.  The discriminators stay plain torch (they are
    not on the set-abstraction path); attribute names / state-dict keys are the reference's."""

    def __init__(self, num_class, num_heading_bin, num_size_cluster, mean_size_arr,
                 input_feature_dim=0, num_proposal=128, vote_factor=1, sampling='vote_fps'):
        super().__init__(num_class, num_heading_bin, num_size_cluster, mean_size_arr,
                         input_feature_dim, num_proposal, vote_factor, sampling)
        self.global_netD1 = nn.Sequential(nn.Conv1d(256, 256, 1), nn.BatchNorm1d(256), nn.ReLU(),
                                          nn.Conv1d(256, 128, 1), nn.BatchNorm1d(128), nn.ReLU())
        self.global_netD2 = nn.Linear(128, 2)
        self.local_netD = nn.Sequential(nn.Conv1d(128, 128, 1), nn.BatchNorm1d(128), nn.ReLU(),
                                        nn.Conv1d(128, 128, 1), nn.BatchNorm1d(128), nn.ReLU(),
                                        nn.Conv1d(128, 1, 1))

    def forward(self, inputs):
        end_points = super().forward(inputs)
        g = self.global_netD1(grad_reverse(end_points['seed_features']))
        end_points['global_d_pred'] = self.global_netD2(torch.mean(g, dim=2))
        local = self.local_netD(grad_reverse(end_points['aggregated_vote_features']))
        end_points['local_d_pred'] = torch.sigmoid(local)
        return end_points
