"""Does the BF16-operand backward of the fused SA blocks train like the reference's arithmetic?

A short optimisation run through the WHOLE backbone (eager steps, torch Adam, lr 1e-3): regress
eight channels of fp2_features onto smooth functions of the seed coordinates, on four rotating
batches of 2 x 8000-point room scenes, from the same initial weights, in four arms:
  fp32        unfused libb2r movers + cuDNN fp32 (the tests' tight arm)
  cudnn_tf32  unfused movers + cuDNN TF32 forward and backward (what the reference runs by default)
  fused       the product: tcgen05 SA blocks (TF32 forward, BF16 backward) + dense tcgen05 FP layers
  fused#2     the product again (run-to-run spread of the atomics-based scatters)
Prints the loss every 10 steps and the mean of the last 10; the product is fine if its curve lies
as close to fp32's as cuDNN-TF32's does.  ("a convergence check for the BF16
backward beyond the one-step rel-L2 bounds".)

    gpurun -- 'python scripts/convergence_check.py > gpurun_out/convergence_check.log 2>&1'
"""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STEPS = int(os.environ.get("B2R_CONV_STEPS", "100"))


def set_arm(arm):
    from backtoreality_b200 import fused_sa
    tf32 = arm == "cudnn_tf32"
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
    fused_sa.ENABLED = arm.startswith("fused")


def target(xyz):
    """(B, n, 3) seed coordinates -> (B, 8, n) values in [0, 1] (fp2_features is post-ReLU)."""
    k = torch.arange(1, 9, device=xyz.device, dtype=torch.float32)[None, :, None]
    s = (xyz[..., 0] + 2.0 * xyz[..., 1] - xyz[..., 2])[:, None, :]
    return 0.5 + 0.5 * torch.sin(k * s)


def train(arm, batches):
    from backtoreality_b200.backbone_module import Pointnet2Backbone
    set_arm(arm)
    torch.manual_seed(7)
    net = Pointnet2Backbone(input_feature_dim=1).cuda().train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    losses = []
    for i in range(STEPS):
        ep = net(batches[i % len(batches)])
        loss = ((ep["fp2_features"][:, :8, :] - target(ep["fp2_xyz"])) ** 2).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return losses


def main():
    from backtoreality_b200 import scenes
    batches = [torch.from_numpy(scenes.batch(100 + 2 * i, 2, 8000, C=1, kind="room", dup=0.2)).cuda()
               for i in range(4)]
    arms = ["fp32", "cudnn_tf32", "fused", "fused#2"]
    curves = {}
    for arm in arms:
        curves[arm] = train(arm, batches)
        print("done", arm, flush=True)
    print("step  " + "  ".join("%-11s" % a for a in arms))
    for s in list(range(0, STEPS, 10)) + [STEPS - 1]:
        print("%4d  " % s + "  ".join("%-11.5f" % curves[a][s] for a in arms))
    tail = {a: sum(curves[a][-10:]) / 10 for a in arms}
    print("mean of the last 10 steps: " + "  ".join("%s %.5f" % kv for kv in tail.items()))
    ref = tail["fp32"]
    print("relative distance to fp32: " + "  ".join(
        "%s %.3f" % (a, abs(tail[a] - ref) / ref) for a in arms[1:]))
    print("first-step loss (same weights, same batch): " + "  ".join(
        "%s %.6f" % (a, curves[a][0]) for a in arms))


if __name__ == "__main__":
    try:
        main()
    except Exception:
        traceback.print_exc()
        sys.exit(1)
