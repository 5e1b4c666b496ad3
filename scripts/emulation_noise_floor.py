"""How far apart can two CORRECT evaluations of one SA block be?  (CPU only, test infrastructure.)

The oracle port with the product's operand rounding (oracle/cpu_modules.emulate_product_operands)
is run twice; the second time the first convolution's output is multiplied by (1 + noise * randn),
noise = 1e-7 / 1e-6: the size of an fp32 summation-order difference.  The block's features and
weight gradients then differ by what tests/test_modules_gpu.py::test_sa_block_vs_operand_rounding_
emulation measures between the fused kernels and the emulation (profiles/r02/emulation_parity.log)
-- i.e. the kernels' residual is indistinguishable from summation order: the next layer's TF32
re-quantisation turns 1e-7 into ~1e-5, and the few ReLU masks / max-pool winners that flip turn
that into 1e-3 .. 1e-2 on the gradients.

    python scripts/emulation_noise_floor.py > profiles/r02/emulation_noise_floor.log
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import pattern_like, rel_l2          # noqa: E402
from backtoreality_b200 import scenes           # noqa: E402
from oracle import cpu_modules as cm            # noqa: E402

CFGS = [dict(N=6000, C=1, npoint=512, radius=0.2, nsample=64, mlp=[1, 64, 64, 128]),
        dict(N=2048, C=128, npoint=1024, radius=0.4, nsample=32, mlp=[128, 128, 128, 256]),
        dict(N=1024, C=256, npoint=512, radius=0.8, nsample=16, mlp=[256, 128, 128, 256])]

for cfg in CFGS:
    torch.manual_seed(11)
    port = cm.emulate_product_operands(cm.SAModuleVotes(
        npoint=cfg["npoint"], radius=cfg["radius"], nsample=cfg["nsample"], mlp=list(cfg["mlp"]),
        normalize_xyz=True).train())
    pc = torch.from_numpy(scenes.batch(31, 2, cfg["N"], C=0, kind="room", dup=0.2))[..., :3].contiguous()
    feats = torch.randn(2, cfg["C"], cfg["N"], generator=torch.Generator().manual_seed(5))

    def run(noise):
        port.zero_grad()
        hook = port.mlp_module.layer0.conv.register_forward_hook(
            lambda m, i, o: o * (1 + noise * torch.randn(o.shape, generator=torch.Generator().manual_seed(9))))
        f = feats.clone().requires_grad_(True)
        _, y, _ = port(pc, f)
        (y * pattern_like(y)).sum().backward()
        hook.remove()
        return y.detach().clone(), {n: p.grad.clone() for n, p in port.named_parameters()}

    y0, g0 = run(0.0)
    for noise in (1e-7, 1e-6):
        y1, g1 = run(noise)
        print("block %-22s noise %.0e  features %.1e  " % (cfg["mlp"], noise, rel_l2(y1.numpy(), y0.numpy()))
              + "  ".join("%s %.1e" % (n.replace("mlp_module.", "").replace(".weight", ""),
                                       rel_l2(g1[n].numpy(), g0[n].numpy()))
                          for n in g0 if "conv" in n))
