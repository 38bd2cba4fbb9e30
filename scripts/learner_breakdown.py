"""Development aid: host-side profile of RNaD.learner_step at cfg2."""
import cProfile, os, pstats, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "r-nad_b200"), REPO]
import torch
import bench
from learn.rnad import RNaD
from nn.net import MLP

depth, a, c, batch = bench.CONFIGS["cfg2"]
dev = torch.device("cuda")
tree = bench.make_tree(depth, a, c); tree.to(dev)
net = MLP(a, 256, device=dev)
trial = RNaD(tree=tree, device=dev, directory_name="prof", batch_size=batch, eta=0.2, lr=1e-3, gamma_averaging=0.01,
             logit_clip=2, net_params={"type": "MLP", "max_actions": a, "width": 256})
trial.net = net; net.train()
trial.net_target, trial.net_reg, trial.net_reg_ = (MLP(a, 256, device=dev) for _ in range(3))
trial.optimizer = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.0, 0.999), eps=1e-8)
def step():
    trial.learner_step(alpha=0.5); trial.total_steps += 1
for _ in range(10): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100): step()
torch.cuda.synchronize()
print("learner_step: %.1f us per call (wall, async)" % ((time.perf_counter() - t0) / 100 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
