"""Development aid: where does the host time of one Episodes.generate call go?"""
import cProfile, os, pstats, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "r-nad_b200"), REPO]
import torch
import bench
from environment.episode import Episodes
from nn.net import MLP

depth, a, c, batch = bench.CONFIGS["cfg2"]
tree = bench.make_tree(depth, a, c); tree.to(torch.device("cuda"))
net = MLP(a, 256, device=torch.device("cuda"))
def step():
    ep = Episodes(tree, batch)
    ep.generate(net)
    return ep
for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
torch.cuda.synchronize()
print("Episodes()+generate: %.1f us per call" % ((time.perf_counter() - t0) / 200 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
