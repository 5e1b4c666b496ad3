"""SharedMLP building blocks with the reference's parameter / state-dict layout.

Mirrors the call surface of /root/reference/detection/Votenet/pointnet2/pytorch_utils.py that the
hot path uses: `SharedMLP` (:11-36), `Conv2d` / `Conv1d` (:122-188), the BatchNorm wrappers
(:39-64) and `BNMomentumScheduler` (:262-296).  Module nesting is what fixes the checkpoint keys
(e.g. `sa1.mlp_module.layer0.conv.weight`, `sa1.mlp_module.layer0.bn.bn.running_mean`,
SURVEY.md appendix B), so it is reproduced exactly; reference checkpoints load unchanged.
The unused Conv3d / FC / preact paths of the reference are not carried over.
"""
import torch.nn as nn


class _BN(nn.Sequential):
    """`<name>bn` child holding the real BatchNorm; weight=1, bias=0 like the reference (:39-47)."""

    def __init__(self, channels, norm_cls, name=""):
        super().__init__()
        self.add_module(name + "bn", norm_cls(channels))
        nn.init.constant_(self[0].weight, 1.0)
        nn.init.constant_(self[0].bias, 0.0)


class BatchNorm1d(_BN):
    def __init__(self, in_size, *, name=""):
        super().__init__(in_size, nn.BatchNorm1d, name)


class BatchNorm2d(_BN):
    def __init__(self, in_size, name=""):
        super().__init__(in_size, nn.BatchNorm2d, name)


class _ConvBlock(nn.Sequential):
    """conv -> (bn) -> (activation); bias only when there is no BN (reference :67-120)."""

    def __init__(self, conv_cls, bn_cls, in_size, out_size, kernel_size, stride, padding,
                 activation, bn, init, bias, name):
        super().__init__()
        conv = conv_cls(in_size, out_size, kernel_size=kernel_size, stride=stride,
                        padding=padding, bias=bias and not bn)
        init(conv.weight)
        if conv.bias is not None:
            nn.init.constant_(conv.bias, 0.0)
        self.add_module(name + "conv", conv)
        if bn:
            self.add_module(name + "bn", bn_cls(out_size))
        if activation is not None:
            self.add_module(name + "activation", activation)


class Conv1d(_ConvBlock):
    def __init__(self, in_size, out_size, *, kernel_size=1, stride=1, padding=0,
                 activation=nn.ReLU(inplace=True), bn=False, init=nn.init.kaiming_normal_,
                 bias=True, preact=False, name=""):
        assert not preact, "preact blocks are not on the hot path"
        super().__init__(nn.Conv1d, BatchNorm1d, in_size, out_size, kernel_size, stride, padding,
                         activation, bn, init, bias, name)


class Conv2d(_ConvBlock):
    def __init__(self, in_size, out_size, *, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0),
                 activation=nn.ReLU(inplace=True), bn=False, init=nn.init.kaiming_normal_,
                 bias=True, preact=False, name=""):
        assert not preact, "preact blocks are not on the hot path"
        super().__init__(nn.Conv2d, BatchNorm2d, in_size, out_size, kernel_size, stride, padding,
                         activation, bn, init, bias, name)


class SharedMLP(nn.Sequential):
    """Stack of 1x1 Conv2d(+BN)+ReLU named layer0, layer1, ... (reference :11-36)."""

    def __init__(self, args, *, bn=False, activation=nn.ReLU(inplace=True), preact=False,
                 first=False, name=""):
        assert not preact, "preact blocks are not on the hot path"
        super().__init__()
        for i in range(len(args) - 1):
            self.add_module(name + "layer{}".format(i),
                            Conv2d(args[i], args[i + 1], bn=bn, activation=activation))

    def channels(self):
        """[C0, C1, ...] as seen by the fused kernels."""
        convs = [blk.conv for blk in self]
        return [convs[0].in_channels] + [c.out_channels for c in convs]


def set_bn_momentum_default(bn_momentum):
    def fn(m):
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
            m.momentum = bn_momentum
    return fn


class BNMomentumScheduler(object):
    """Sets BatchNorm momentum from `bn_lambda(epoch)` (reference :271-296)."""

    def __init__(self, model, bn_lambda, last_epoch=-1, setter=set_bn_momentum_default):
        if not isinstance(model, nn.Module):
            raise RuntimeError("Class '{}' is not a PyTorch nn Module".format(type(model).__name__))
        self.model = model
        self.setter = setter
        self.lmbd = bn_lambda
        self.step(last_epoch + 1)
        self.last_epoch = last_epoch

    def step(self, epoch=None):
        if epoch is None:
            epoch = self.last_epoch + 1
        self.last_epoch = epoch
        self.model.apply(self.setter(self.lmbd(epoch)))
