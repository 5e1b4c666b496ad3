"""Adam over one flat parameter buffer (csrc/adam.cu, include/b2r.h: b2r_adam_flat_step).

Mirrors `optim.Adam(net.parameters(), lr=..., weight_decay=...)` of the reference's training
scripts (/root/reference/detection/Votenet/train_Votenet_FSB.py:176-181) for the captured step:
parameters become views of ONE contiguous fp32 buffer, the step's gradients are packed into a
second one (the buffer the NCCL all-reduce of dist_utils.FlatGradBucket already works on) and the
update is a single streaming kernel whose step counter lives on the device, so it replays from a
CUDA graph.  Same arithmetic as torch.optim.Adam (amsgrad=False, maximize=False); checked against
it in tests/test_modules_gpu.py.

    opt = FlatAdam(net.parameters(), lr=1e-3)
    loss.backward(); opt.step()             # packs p.grad of every parameter, then updates
    opt.step(grads=list_of_tensors)         # or: gradients given explicitly (p.grad = None scheme)
"""
import torch
import torch.distributed as dist

from . import _ext, _lib


class FlatAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("FlatAdam: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam runs on CUDA only (there is no CPU fallback)")
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("FlatAdam: parameters must be fp32 tensors on one device")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        # 16-byte aligned slices: every parameter starts at a multiple of 4 floats
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.numel = n
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.state = torch.zeros(int(_lib.lib().b2r_adam_state_bytes()), dtype=torch.uint8, device=dev)
        self.g_views = []
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = self.flat_p[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view                       # the module's Parameter now lives in the flat buffer
                self.g_views.append(self.flat_g[off:off + p.numel()].view_as(p))

    # ---- overlapping the gradient all-reduce with the rest of backward (world > 1) --------------
    def set_late(self, n_late):
        """The first `n_late` parameters are the ones whose gradients arrive LAST in backward (the
        first layers of the network: backward runs the forward order in reverse).  `reduce_early`
        may then be called from a backward hook placed where everything after them is done."""
        self.n_late = int(n_late)
        self.off_late = self.offsets[self.n_late] if self.n_late < len(self.params) else self.numel
        self.comm = torch.cuda.Stream(device=self.flat_p.device)
        self._early_done = False

    def reduce_early(self):
        """pack + all-reduce the gradients of params[n_late:] on a side stream, while backward goes
        on with the first layers.  Call when those gradients exist (tensor hook on an activation
        between the two parameter groups).  No-op without a process group."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        early = self.params[self.n_late:]
        if any(p.grad is None for p in early):
            return                      # not all there (e.g. a frozen branch): step() does everything
        torch._foreach_copy_(self.g_views[self.n_late:], [p.grad for p in early])
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            dist.all_reduce(self.flat_g[self.off_late:])
        self._early_done = True

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None

    def pack(self, grads=None):
        """gradients of this step -> the flat buffer (one multi-tensor copy); parameters without a
        gradient contribute zeros"""
        if grads is None:
            grads = [p.grad for p in self.params]
        have = [(v, g) for v, g in zip(self.g_views, grads) if g is not None]
        if len(have) != len(self.g_views):
            self.flat_g.zero_()
        torch._foreach_copy_([v for v, _ in have], [g for _, g in have])

    def step(self, grads=None, allreduce=True):
        """pack -> (world > 1: one NCCL sum all-reduce of the flat gradient) -> one Adam kernel"""
        scale = 1.0
        multi = allreduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if multi and getattr(self, "_early_done", False):
            # params[n_late:] are packed and their all-reduce is in flight on the side stream
            self._early_done = False
            if grads is None:
                grads = [p.grad for p in self.params]
            late = [(v, g) for v, g in zip(self.g_views[:self.n_late], grads[:self.n_late]) if g is not None]
            if len(late) != self.n_late:
                self.flat_g[:self.off_late].zero_()
            torch._foreach_copy_([v for v, _ in late], [g for _, g in late])
            if self.off_late > 0:
                dist.all_reduce(self.flat_g[:self.off_late])
            torch.cuda.current_stream().wait_stream(self.comm)
        else:
            self.pack(grads)
            if multi:
                dist.all_reduce(self.flat_g)
        if multi:
            scale = 1.0 / dist.get_world_size()     # averaged inside the Adam kernel (DDP semantics)
        with _ext._on_device(self.flat_p):
            _lib.check(_lib.lib().b2r_adam_flat_step(
                self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
                self.exp_avg_sq.data_ptr(), self.numel, self.state.data_ptr(), self.lr, self.betas[0],
                self.betas[1], self.eps, self.weight_decay, scale, _ext._stream()), "adam_flat_step")
        # the kernel wrote through raw pointers: tell autograd / the weight-image caches
        torch.autograd.graph.increment_version(self.params)

    def state_dict(self):
        return {"flat_p": self.flat_p.clone(), "exp_avg": self.exp_avg.clone(),
                "exp_avg_sq": self.exp_avg_sq.clone(), "state": self.state.clone(),
                "lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd):
        self.flat_p.copy_(sd["flat_p"]); self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"]); self.state.copy_(sd["state"])
        self.lr, self.betas, self.eps, self.weight_decay = sd["lr"], sd["betas"], sd["eps"], sd["weight_decay"]
