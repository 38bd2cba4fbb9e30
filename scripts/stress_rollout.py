"""Development aid: repeat the fused rollout many times and look for gross net-output errors (ordering races).

The recorded observations go through a plain fp32 torch forward on the GPU; the tensor-core engines differ from it by
< 1e-2, a mis-ordered tensor-memory access by much more.  usage: stress_rollout.py GOLDEN_NAME BATCH ITERS [precision]
"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "r-nad_b200"), REPO, os.path.join(REPO, "tests")]
import numpy as np, torch
from helpers import gross_rollout_errors, tree_from_golden
from environment.episode import Episodes
from nn.net import MLP

name, B, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
prec = sys.argv[4] if len(sys.argv) > 4 else "tf32x2"
data = np.load(os.path.join(REPO, "tests", "golden", name + ".npz"))
g = {k: data[k] for k in data.files}
dev = torch.device("cuda")
tree = tree_from_golden(g, "cuda")
a = tree.max_actions
torch.manual_seed(7)
net = MLP(a, 256, device=dev)
with torch.no_grad():
    for p in net.parameters():
        p.mul_(2.0)
bad_runs = 0
for it in range(iters):
    torch.manual_seed(1000 + it)
    ep = Episodes(tree, B)
    ep.generate(net, precision=prec)
    bad_v, bad_p, val = gross_rollout_errors(ep, net)
    valid = ep.indices[: ep.t_eff + 1] != 0
    nv, npol = int(bad_v.sum()), int(bad_p.sum())
    if nv or npol:
        bad_runs += 1
        print(f"run {it}: {nv} wrong values, {npol} wrong policies (of {int(valid.sum())} valid slots)")
        for nm, bad in (("value", bad_v), ("policy", bad_p)):
            if not int(bad.sum()):
                continue
            tt, gg = torch.nonzero(bad, as_tuple=True)
            for t_ in torch.unique(tt).tolist():
                games = gg[tt == t_]
                tiles = torch.unique(games // 128)
                print(f"   {nm}: half-move {t_}: {len(games)} games in tiles {tiles[:10].tolist()} "
                      f"(pair, side: {[(x // 2, x % 2) for x in tiles[:6].tolist()]}), lanes {torch.unique(games % 128)[:6].tolist()}..., "
                      f"{len(torch.unique(games % 128))} distinct; err {float((ep.values[t_, games] - val[t_, games]).abs().max()):.3f}")
print(f"{bad_runs} of {iters} runs with gross errors ({name}, batch {B}, {prec})")
