"""2-rank NCCL check of flat_adam.FlatAdam with the overlapped early bucket (torchrun, 2 GPUs):
two ranks train a small two-stage model on disjoint halves of a batch, the early bucket is reduced
from a tensor hook in the middle of backward; the result must equal single-process Adam on the
averaged gradients.  Prints 'dist_flat_adam ok' on rank 0."""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backtoreality_b200.flat_adam import FlatAdam

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)


def make():
    torch.manual_seed(0)
    a = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU()).to(dev)   # "late" stage (first layers)
    b = torch.nn.Sequential(torch.nn.Linear(32, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4)).to(dev)
    return a, b


g = torch.Generator().manual_seed(1)
data = torch.randn(3, 8, 16, generator=g).to(dev)
a, b = make()
params = list(a.parameters()) + list(b.parameters())
opt = FlatAdam(params, lr=1e-2)
opt.set_late(len(list(a.parameters())))
ra, rb = make()
ropt = torch.optim.Adam(list(ra.parameters()) + list(rb.parameters()), lr=1e-2)
for it in range(3):
    x = data[it]
    # reference: gradients of every rank's half, averaged (DDP semantics)
    ropt.zero_grad()
    for r in range(world):
        xs = x[r * 4:(r + 1) * 4]
        (rb(ra(xs)).square().mean() / world).backward()
    ropt.step()
    opt.zero_grad()
    h = a(x[rank * 4:(rank + 1) * 4])
    h.register_hook(lambda gr: opt.reduce_early())
    b(h).square().mean().backward()
    assert opt._early_done
    opt.step()
err = max(float((p - q).abs().max()) for p, q in zip(params, list(ra.parameters()) + list(rb.parameters())))
assert err < 1e-5, err
if rank == 0:
    print("dist_flat_adam ok, max |dp| = %.2e" % err, flush=True)
torch.cuda.synchronize()
os._exit(0)
