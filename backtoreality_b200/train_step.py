"""One training step as a replayable CUDA graph.

The hot path issues ~450 small launches per step (libb2r kernels through ctypes, cuDNN for the
FP / vote heads, optimizer); at ~9 ms of GPU work the Python launch overhead leaves the GPU idle
~20 % of the step.  Everything in the step is shape-static and sync-free (no `.item()`, no host
reads), so forward + backward + gradient all-reduce + optimizer are captured ONCE into a CUDA
graph and replayed: one launch per step, no gaps.

Values that are baked in at capture time: tensor addresses (the input is a static buffer that
`__call__` copies into), BatchNorm momentum (a Python float the reference's BNMomentumScheduler
changes once per epoch, pytorch_utils.py:262-296 -> call `recapture()` after changing it) and the
learning rate unless it is a tensor.
"""
import torch


class CapturedTrainStep:
    """`step_fn(static_input) -> loss tensor` runs fwd + bwd + all-reduce + optimizer, eagerly.
    This wraps it: `loss = captured(batch)` copies `batch` into the static input and replays."""

    def __init__(self, step_fn, example_input, warmup=3, after_warmup_step=None):
        self.step_fn = step_fn
        self.after_warmup_step = after_warmup_step   # e.g. the eager optimizer / gradient reset
        self.static_in = example_input.clone()
        self.warmup = warmup
        self.graph = None
        self.static_loss = None
        self.launches_per_step = None
        self.recapture()

    def recapture(self):
        from . import _ext
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self.step_fn(self.static_in)
                if self.after_warmup_step is not None:
                    self.after_warmup_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        saved_ops = set(_ext.TIME_OPS)
        _ext.TIME_OPS.clear()          # timing events cannot be recorded inside a capture
        l0 = _ext.LAUNCHES
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss = self.step_fn(self.static_in)
            self.graph, self.static_loss = g, loss
            self.launches_per_step = _ext.LAUNCHES - l0
        finally:
            _ext.TIME_OPS.update(saved_ops)

    def __call__(self, batch, non_blocking=True):
        self.static_in.copy_(batch, non_blocking=non_blocking)
        self.graph.replay()
        return self.static_loss
