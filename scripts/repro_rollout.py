"""Development aid: launch the fused rollout again and again with ONE seed and report how launches differ from the first.
usage: repro_rollout.py GOLDEN_NAME BATCH ITERS [precision]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "r-nad_b200"), REPO, os.path.join(REPO, "tests")]
import numpy as np, torch
from helpers import tree_from_golden
from environment.episode import Episodes
from nn.net import MLP

name, B, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
prec = sys.argv[4] if len(sys.argv) > 4 else "tf32x2"
data = np.load(os.path.join(REPO, "tests", "golden", name + ".npz"))
g = {k: data[k] for k in data.files}
tree = tree_from_golden(g, "cuda")
torch.manual_seed(7)
net = MLP(tree.max_actions, 256, device=torch.device("cuda"))
with torch.no_grad():
    for p in net.parameters():
        p.mul_(2.0)
first = None
for it in range(iters):
    ep = Episodes(tree, B)
    ep.states.seed = 1234
    ep.generate(net, precision=prec)
    T = ep.t_eff + 1
    cur = {k: ep.full(k)[:T].clone() for k in ("indices", "actions", "policy", "values")}
    if first is None:
        first = cur
        continue
    same_path = torch.equal(cur["indices"], first["indices"]) and torch.equal(cur["actions"], first["actions"])
    dv = (cur["values"] - first["values"]).abs()
    dp = (cur["policy"] - first["policy"]).abs().max(-1).values
    nv, npol = int((dv > 0).sum()), int((dp > 0).sum())
    if nv or npol or not same_path:
        tt, gg = torch.nonzero((dv > 0) | (dp > 0), as_tuple=True)
        tiles = torch.unique(gg // 128)
        print(f"launch {it}: same trajectories {same_path}; {nv} values differ (max {float(dv.max()):.3e}), {npol} policies differ "
              f"(max {float(dp.max()):.3e}); half-moves {torch.unique(tt).tolist()}, {len(tiles)} tiles {tiles[:8].tolist()}, "
              f"lanes {torch.unique(gg % 128)[:8].tolist()} ({len(torch.unique(gg % 128))} distinct)")
print("done", name, B, prec)
