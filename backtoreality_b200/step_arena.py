"""One zero-filled arena per training step.

The fused SA blocks and the dense MLPs accumulate into zero-initialised buffers (BatchNorm
statistics, weight gradients, scatter targets): ~40 `torch.zeros` per VoteNet step, each a fill
kernel sitting in front of its consumer on the step's critical path.  Inside a captured step
(train_step.CapturedTrainStep / PipelinedTrainStep) they are carved from ONE buffer that is
cleared by a single memset at the top of the step.

Only the step wrappers call `begin()`: the buffers handed out are valid until the next `begin()`
on the same device, which is exactly the lifetime of a step's activations and gradients THERE
(gradients are consumed by the optimizer inside the step).  Everywhere else `zeros()` is
`torch.zeros`.  B2R_ZERO_ARENA=0 disables it.
"""
import os

import torch

ENABLED = os.environ.get("B2R_ZERO_ARENA", "1") not in ("0", "")
_ARENAS = {}      # device -> _Arena


class _Arena:
    def __init__(self, device, nbytes):
        self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self.off = 0
        self.active = False
        self.want = 0          # bytes asked for during the current step (sizing the next one)


def begin(device, min_bytes=48 << 20):
    """Start a step on `device`: one memset clears every buffer the step will take."""
    if not ENABLED:
        return
    a = _ARENAS.get(device)
    capturing = torch.device(device).type == "cuda" and torch.cuda.is_current_stream_capturing()
    if a is None or (a.want > a.buf.numel() and not capturing):
        a = _ARENAS[device] = _Arena(device, max(min_bytes, int((a.want if a else 0) * 1.25)))
    else:
        a.buf.zero_()
    a.off, a.want, a.active = 0, 0, True


def end(device):
    a = _ARENAS.get(device)
    if a is not None:
        a.active = False


def zeros(shape, dtype, device):
    a = _ARENAS.get(device) if ENABLED else None
    if a is None or not a.active:
        return torch.zeros(shape, dtype=dtype, device=device)
    numel = 1
    for s in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)):
        numel *= int(s)
    nbytes = (numel * torch.empty((), dtype=dtype).element_size() + 255) & ~255
    a.want += nbytes
    if a.off + nbytes > a.buf.numel():
        return torch.zeros(shape, dtype=dtype, device=device)     # overflow: plain fill this time
    out = a.buf[a.off:a.off + nbytes].view(dtype)[:numel].view(shape)
    a.off += nbytes
    return out


def zeros_like(t):
    return zeros(tuple(t.shape), t.dtype, t.device)
