"""Development aid: back-to-back time of the fused rollout at cfg2 for the library in $RNAD_B200_LIB (no result checks)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "r-nad_b200"), REPO]
import torch
import bench
from nn.net import MLP

depth, a, c, batch = bench.CONFIGS["cfg2"]
tree = bench.make_tree(depth, a, c); tree.to(torch.device("cuda"))
torch.manual_seed(1234)
net = MLP(a, 256, device=torch.device("cuda"))
out = []
for prec in sys.argv[1:] or ["tf32x2", "f16x2"]:
    r = bench.RolloutRunner(tree, net, batch, prec)
    for _ in range(20):
        r.launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2000):
        r.launch()
    e1.record(); e1.synchronize()
    out.append(f"{prec} {e0.elapsed_time(e1) / 2000 * 1e3:.2f} us")
print(os.path.basename(os.environ.get("RNAD_B200_LIB", "default")), " ".join(out), flush=True)
