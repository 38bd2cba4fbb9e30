"""Development aid: how does tcgen05 kind::tf32 treat an fp32 A operand read from tensor memory - truncate or round?"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(REPO, "r-nad_b200"), REPO, os.path.join(REPO, "tests")]
import torch
from oracle import rnad_oracle as orc
from environment.episode import Episodes
from environment.tree import Tree
from environment.fast_tree import depth_jitter
from test_gpu_env_rollout import wide_net

tree = Tree(device=torch.device("cuda"), max_actions=3, max_transitions=2, depth_bound=6, transition_threshold=0.3)
tree.generate_fast(seed=3, child_spec=depth_jitter(0.4))
net, w = wide_net(3, 5, "cuda")
ep = Episodes(tree, 30000); ep.generate(net)
obs = ep.observations.cpu().reshape(-1, 18); val = ep.values.cpu().reshape(-1).double()
x = orc.tf32_rna(obs).double()
def value(mode):
    w0 = orc.tf32_rna(w["value_fc0.weight"]).double(); b0 = orc.tf32_rna(w["value_fc0.bias"]).double()
    h = torch.relu(x @ w0.T + b0)
    h32 = h.float()
    if mode == "trunc": h = orc.tf32_trunc(h32).double()
    elif mode == "rna": h = orc.tf32_rna(h32).double()
    elif mode == "fp32": h = h32.double()
    w1 = orc.tf32_rna(w["value_fc1.weight"]).double()
    return (h @ w1.T + w["value_fc1.bias"].double())[:, 0]
for mode in ("trunc", "rna", "fp32"):
    err = (value(mode) - val).abs()
    print(f"h as {mode:6s}: max err {float(err.max()):.3e} mean {float(err.mean()):.3e}")
