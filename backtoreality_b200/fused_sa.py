"""Fused SharedMLP block of a set-abstraction layer on tcgen05 tensor cores (host side).

Drives csrc/mlp.cu through the C ABI (`b2r_mlp_pack_weight`, `b2r_sa_layer_fwd`,
`b2r_bn_finalize`, `b2r_pool_finalize`, include/b2r.h) to compute what the reference computes in
PointnetSAModuleVotes.forward (pointnet2_modules.py:245-267) with
QueryAndGroup -> SharedMLP(3 x [Conv2d 1x1 -> BatchNorm2d -> ReLU]) -> max_pool2d:

    layer 0   gather (idx) + relative/normalised xyz  -> TF32 GEMM -> raw z0 + batch statistics
    layer i   relu(bn(z_{i-1})) applied on load       -> TF32 GEMM -> raw z_i + batch statistics
    last      ... -> TF32 GEMM -> statistics + max/min over nsample (no (B,C,np,ns) tensor)
    finalize  BatchNorm scale/shift (+ running-stat update) and relu(bn(max|min)) -> (B,C,np)

Training-mode BatchNorm uses biased batch statistics over all B*npoint*nsample positions of the
local batch, updates running_mean / running_var (unbiased) with the module's momentum and bumps
num_batches_tracked, exactly like nn.BatchNorm2d (reference pytorch_utils.py:55-58).
Math: TF32 operands (round-to-nearest), FP32 accumulate -- the precision the reference itself
runs at by default (cuDNN TF32 convolutions on sm_80+).
"""
import ctypes

import torch

from . import _ext, _lib

_vp = ctypes.c_void_p


def _ptr(t):
    return _vp(t.data_ptr()) if t is not None else None


def to_point_major(features):
    """(B,C,N) -> (B,N,C) contiguous."""
    B, C, N = features.shape
    out = torch.empty((B, N, C), dtype=torch.float32, device=features.device)
    _lib.check(_lib.lib().b2r_to_point_major(_ptr(features), B, C, N, _ptr(out), _ext._stream()),
               "to_point_major")
    _ext.LAUNCHES += 1
    return out


def pack_weight(weight, gather):
    """Conv weight (Cout,Cin,1,1) -> packed TF32 swizzled image (1-D float tensor)."""
    Cout, Cin = weight.shape[0], weight.shape[1]
    w = weight.detach().reshape(Cout, Cin).contiguous()
    nbytes = _lib.lib().b2r_mlp_weight_image_bytes(Cout, Cin, 1 if gather else 0)
    image = torch.empty(nbytes // 4, dtype=torch.float32, device=weight.device)
    _lib.check(_lib.lib().b2r_mlp_pack_weight(_ptr(w), Cout, Cin, 1 if gather else 0, _ptr(image),
                                              _ext._stream()), "mlp_pack_weight")
    _ext.LAUNCHES += 1
    return image


def supported(mlp_module, xyz, features, idx, pooling="max"):
    """True when the fused kernels cover this block (otherwise callers use the unfused path)."""
    if pooling != "max" or not xyz.is_cuda:
        return False
    B, NP, NS = idx.shape
    if NS not in (16, 32, 64) or (B * NP * NS) % 128 != 0:
        return False
    blocks = list(mlp_module)
    if len(blocks) == 0:
        return False
    for i, blk in enumerate(blocks):
        if not hasattr(blk, "bn") or blk.conv.bias is not None:
            return False
        bn = blk.bn.bn
        if bn.momentum is None or not bn.track_running_stats or not bn.affine:
            return False
        cout, cin = blk.conv.out_channels, blk.conv.in_channels
        if cout > 256 or (i > 0 and cin % 4 != 0):
            return False
        smem = _layer_smem(cin, cout, gather=(i == 0), nt=64)
        if smem > 227 * 1024:
            return False
        if i == len(blocks) - 1 and cout <= 128 and \
                _layer_smem(cin, cout, gather=(i == 0), nt=128) > 227 * 1024:
            return False
    return True


def _layer_smem(cin, cout, gather, nt):
    kp = (((cin - 3 + 3) & ~3) + 4) if gather else ((cin + 3) & ~3)
    ka = (kp + 31) >> 5
    cout_pad = (cout + 127) & ~127
    return 1024 + cout_pad * ka * 128 + nt * ka * 128 + 2 * kp * 4 + nt * 4 + 48


def sa_mlp_forward(xyz, new_xyz, feat_t, idx, radius, normalize_xyz, mlp_module, training,
                   want_point_major=True, save=None):
    """Run the fused block.

    xyz (B,N,3), new_xyz (B,NP,3), feat_t (B,N,C) POINT-major features or None, idx (B,NP,NS).
    Returns (out_cm (B,Cl,NP), out_pm (B,NP,Cl) or None).  When `save` is a dict it receives what
    a backward pass needs (raw z per layer, BN mean/invstd/scale/shift, pooled max/min/arg).
    """
    lib = _lib.lib()
    st = _ext._stream()
    dev = xyz.device
    B, N = xyz.shape[0], xyz.shape[1]
    NP, NS = idx.shape[1], idx.shape[2]
    M = B * NP * NS
    blocks = list(mlp_module)
    L = len(blocks)
    z_prev = scale = shift = None
    zs, bn_saved = [], []
    zmax = zmin = amax = amin = None
    for i, blk in enumerate(blocks):
        conv, bn = blk.conv, blk.bn.bn
        Cin, Cout = conv.in_channels, conv.out_channels
        last = i == L - 1
        image = pack_weight(conv.weight, gather=(i == 0))
        stats = torch.zeros((2, Cout), dtype=torch.float64, device=dev) if training else None
        d = _lib.SaLayer()
        d.B, d.N, d.NP, d.NS, d.Cin, d.Cout = B, N, NP, NS, Cin, Cout
        d.mode = 0 if i == 0 else 1
        d.epilogue = 1 if last else 0
        if i == 0:
            d.xyz, d.new_xyz, d.feat_t, d.idx = _ptr(xyz), _ptr(new_xyz), _ptr(feat_t), _ptr(idx)
            d.radius, d.normalize_xyz = float(radius), 1 if normalize_xyz else 0
        else:
            d.z_prev, d.scale_prev, d.shift_prev = _ptr(z_prev), _ptr(scale), _ptr(shift)
        d.w_image = _ptr(image)
        z = None
        if last:
            zmax = torch.empty((B * NP, Cout), dtype=torch.float32, device=dev)
            zmin = torch.empty_like(zmax)
            amax = torch.empty((B * NP, Cout), dtype=torch.int32, device=dev)
            amin = torch.empty_like(amax)
            d.zmax, d.zmin, d.amax, d.amin = _ptr(zmax), _ptr(zmin), _ptr(amax), _ptr(amin)
        else:
            z = torch.empty((M, Cout), dtype=torch.float32, device=dev)
            d.z = _ptr(z)
        d.stats = _ptr(stats)
        _lib.check(lib.b2r_sa_layer_fwd(ctypes.byref(d), st), "sa_layer_fwd")
        _ext.LAUNCHES += 1
        # BatchNorm scale / shift of THIS layer (applied by the next layer's prologue / finalize)
        if training:
            scale = torch.empty(Cout, dtype=torch.float32, device=dev)
            shift = torch.empty_like(scale)
            mean = torch.empty_like(scale)
            invstd = torch.empty_like(scale)
            _lib.check(lib.b2r_bn_finalize(_ptr(stats), Cout, float(M), _ptr(bn.weight),
                                           _ptr(bn.bias), float(bn.eps), float(bn.momentum),
                                           _ptr(bn.running_mean), _ptr(bn.running_var),
                                           _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd), st),
                       "bn_finalize")
            _ext.LAUNCHES += 1
            bn.num_batches_tracked.add_(1)
        else:
            invstd = torch.rsqrt(bn.running_var + bn.eps)
            mean = bn.running_mean
            scale = (bn.weight * invstd).contiguous()
            shift = (bn.bias - bn.running_mean * scale).contiguous()
        zs.append(z)
        bn_saved.append((mean, invstd, scale, shift))
        z_prev = z
    Cl = blocks[-1].conv.out_channels
    out_cm = torch.empty((B, Cl, NP), dtype=torch.float32, device=dev)
    out_pm = torch.empty((B, NP, Cl), dtype=torch.float32, device=dev) if want_point_major else None
    _lib.check(lib.b2r_pool_finalize(_ptr(zmax), _ptr(zmin), _ptr(scale), _ptr(shift), B, NP, Cl,
                                     _ptr(out_cm), _ptr(out_pm), st), "pool_finalize")
    _ext.LAUNCHES += 1
    if save is not None:
        save.update(zs=zs, bn=bn_saved, zmax=zmax, zmin=zmin, amax=amax, amin=amin)
    return out_cm, out_pm
