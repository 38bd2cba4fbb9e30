"""Development aid: where does a tf32x2 rollout differ from the tf32-aware oracle?  usage: debug_rollout.py A C depth B"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(REPO, "r-nad_b200"), REPO]
import numpy as np, torch
import bench
from oracle import rnad_oracle as orc
from environment.episode import Episodes
from nn.net import MLP

a, c, depth, B = (int(x) for x in sys.argv[1:5])
prec = sys.argv[5] if len(sys.argv) > 5 else "tf32x2"
tree = bench.make_tree(depth, a, c)
tables = {"expected_value": tree.expected_value_tensor.clone(), "legal": tree.legal_tensor.clone()}
tree.to(torch.device("cuda"))
torch.manual_seed(3)
net = MLP(a, 256, device=torch.device("cuda"))
w = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
ep = Episodes(tree, B)
ep.generate(net, precision=prec)
T = ep.t_eff + 1
for s in range(T):
    idx = ep.indices[s].cpu()
    obs = orc.observe(tables["expected_value"], tables["legal"], idx, s & 1)
    ok_obs = (ep.observations[s].cpu() == obs).flatten(1).all(1)
    _, pol, val, _, _, _ = orc.mlp_forward_tc(w, obs.reshape(B, -1), "tf32" if prec == "tf32x2" else "fp32")
    bad = ((ep.values[s].cpu().double() - val[:, 0]).abs() > 1e-4) | ((ep.policy[s].cpu().double() - pol).abs().max(1).values > 1e-4)
    nb = int(bad.sum())
    print(f"half-move {s}: obs mismatches {int((~ok_obs).sum())}, net-output mismatches {nb}", end="")
    if nb:
        g = torch.nonzero(bad).flatten()
        tiles = torch.unique(g // 128)
        for tl in tiles[:8].tolist():
            lanes = (g[(g // 128) == tl] % 128).tolist()
            print(f"\n   tile {tl} (pair {tl // 2} side {tl % 2}): {len(lanes)} lanes, warps {sorted(set(l // 32 for l in lanes))}, first {lanes[:4]}", end="")
        gg = g[:3].tolist() + g[-2:].tolist()
        for q in gg:
            prev = q - 2 * 148 * 128
            print(f"\n   game {q}: value gpu {float(ep.values[s, q]):+.5f} want {float(val[q, 0]):+.5f}; same lane, previous pair: "
                  + " ".join(f"t{tt}:{float(ep.values[tt, prev]):+.5f}" for tt in range(T)) + f"  policy gpu {ep.policy[s, q].cpu().tolist()} want {pol[q].tolist()}", end="")
        print(f"\n   games {g[:6].tolist()}..., tiles {tiles[:12].tolist()} ({len(tiles)} tiles), lanes {torch.unique(g % 128)[:8].tolist()}"
              f"; max err {float((ep.values[s].cpu().double() - val[:, 0]).abs().max()):.3e}", end="")
    print()
