"""Find host synchronisations / slow host code in one eager VoteNet step (debug helper)."""
import cProfile, pstats, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from backtoreality_b200 import scenes
from backtoreality_b200.votenet import VoteNet
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = VoteNet(22, 1, 22, np.ones((22, 3), np.float32), input_feature_dim=1, num_proposal=256).to(dev).train()
pc = torch.from_numpy(scenes.batch(0, 8, 40000, C=1, kind="room", dup=0.2)).to(dev)
def step():
    for p in net.parameters():
        p.grad = None
    ep = net({"point_clouds": pc})
    loss = (ep["proposal_scores_raw"] ** 2).mean() + ((ep["vote_xyz"] - ep["seed_xyz"]) ** 2).mean()
    loss.backward()
    return loss
for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print("eager step %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
torch.cuda.set_sync_debug_mode("warn")
step()
torch.cuda.set_sync_debug_mode("default")
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
