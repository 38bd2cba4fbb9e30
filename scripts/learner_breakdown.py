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
# CPU time of one call when the GPU is idle (each call followed by a synchronize): the serial host path of a step
loss_host = torch.empty(2).pin_memory()
ts = {"call": 0.0, "copy": 0.0, "sync": 0.0}
for _ in range(200):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter()
    loss_host.copy_(trial.last_losses, non_blocking=True); t2 = time.perf_counter()
    torch.cuda.current_stream().synchronize(); t3 = time.perf_counter()
    ts["call"] += t1 - t0; ts["copy"] += t2 - t1; ts["sync"] += t3 - t2
print({k: round(v / 200 * 1e6, 1) for k, v in ts.items()}, "us per step (call = host time of learner_step, sync includes the GPU's 0.23 ms)")
