"""eager vs eager vs graph vs graph loss trajectories (same seeds)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backtoreality_b200 import scenes
from backtoreality_b200.train_step import CapturedTrainStep
from backtoreality_b200.votenet import VoteNet
cuda = torch.device("cuda:0")

def make():
    torch.manual_seed(5)
    net = VoteNet(4, 1, 4, np.ones((4, 3), np.float32), input_feature_dim=1, num_proposal=64,
                  vote_factor=1, sampling="vote_fps").to(cuda).train()
    opt = torch.optim.SGD(net.parameters(), lr=1e-2)
    def step(pc):
        ep = net({"point_clouds": pc})
        loss = (ep["proposal_scores_raw"] ** 2).mean() + ((ep["vote_xyz"] - ep["seed_xyz"]) ** 2).mean()
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=False)
        return loss.detach()
    return net, step

batches = [torch.from_numpy(scenes.batch(300 + 2 * i, 2, 8192, C=1, kind="room", dup=0.2)).to(cuda) for i in range(4)]
for name in ("eager", "eager", "graph", "graph"):
    net, step = make()
    if name == "eager":
        for i in range(3):
            step(batches[0])
        out = [float(step(b)) for b in batches]
    else:
        cap = CapturedTrainStep(step, batches[0], warmup=3)
        out = [float(cap(b)) for b in batches]
    print(name, ["%.6f" % v for v in out])
# forward-only determinism of the net on one batch: eager twice
net, step = make()
with torch.no_grad():
    a = net({"point_clouds": batches[1]})["proposal_scores_raw"].clone()
    b = net({"point_clouds": batches[1]})["proposal_scores_raw"].clone()
print("fwd twice max diff", float((a - b).abs().max()))
