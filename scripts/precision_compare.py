#!/usr/bin/env python
"""Error of every rollout engine against a float64 forward of the recorded observations, on the cfg2 tree and batch
(GPU box):  python scripts/precision_compare.py  -> one line per engine (policy / value: max and mean absolute error)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "r-nad_b200"), REPO]
import torch  # noqa: E402

import bench  # noqa: E402
from environment.episode import Episodes  # noqa: E402
from nn.net import MLP  # noqa: E402

dev = torch.device("cuda:0")
tree = bench.make_tree(4, 3, 2, seed=0)
tree.to(dev)
for scale in (1.0, 2.0):
    torch.manual_seed(1234)
    net = MLP(3, 256, device=dev)
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(scale)
    net64 = MLP(3, 256, device=dev).double()
    net64.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
    for precision in ("fp32", "tf32", "tf32x2", "f16x2"):
        torch.manual_seed(7)
        ep = Episodes(tree, 65536)
        ep.generate(net, precision=precision)
        obs = ep.observations.double()
        t, b = obs.shape[:2]
        with torch.no_grad():
            _, _, policy, value = net64.forward(obs.reshape(t * b, *obs.shape[2:]))[:4] if False else (None, None, None, None)
            x = obs.reshape(t * b, -1)
            sd = net64.state_dict()
            hv = torch.relu(x @ sd["value_fc0.weight"].T + sd["value_fc0.bias"])
            value = (hv @ sd["value_fc1.weight"].T + sd["value_fc1.bias"])[:, 0]
            hp = torch.relu(x @ sd["policy_fc0.weight"].T + sd["policy_fc0.bias"])
            logits = hp @ sd["policy_fc1.weight"].T + sd["policy_fc1.bias"]
            mask = ep.masks.reshape(t * b, -1) != 0
            e = torch.where(mask, torch.exp(logits), torch.zeros_like(logits))
            policy = e / e.sum(-1, keepdim=True).clamp_min(1e-12)
        valid = (ep.indices.reshape(-1) != 0)
        dp = (ep.policy.reshape(t * b, -1).double() - policy).abs()[valid]
        dv = (ep.values.reshape(-1).double() - value).abs()[valid]
        print(f"weights x{scale}  {precision:7s} policy max {dp.max().item():.2e} mean {dp.mean().item():.2e}   "
              f"value max {dv.max().item():.2e} mean {dv.mean().item():.2e}", flush=True)
