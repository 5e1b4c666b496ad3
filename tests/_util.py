"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def weight_checksum(module):
    return float(sum(p.detach().double().abs().sum().cpu() for p in module.parameters()))


def sub(t, n=4096):
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step].cpu().numpy().copy()


def rel_l2(got, want):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    return float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30))


def pattern_like(t):
    """Fixed pseudo-random loss weights in [-1, 1] (float64 sine hash -> fp32).

    NOT a linear ramp: a ramp is almost constant along the position axis of every channel, and
    BatchNorm's backward removes exactly that constant, so the surviving gradient is pure
    cancellation residue (measured: the fp32 CPU reference then disagrees with ITSELF by 6e-3
    between 1 and 8 threads)."""
    i = torch.arange(t.numel(), dtype=torch.float64)
    return torch.sin(i * 12.9898 + 0.5 * torch.cos(i * 0.618)).float().reshape(t.shape).to(t.device)
