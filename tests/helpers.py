"""Shared helpers for the parity tests (golden fixture access)."""
import numpy as np
import torch

NET_KEYS = ("value_fc0.weight", "value_fc0.bias", "value_fc1.weight", "value_fc1.bias",
            "policy_fc0.weight", "policy_fc0.bias", "policy_fc1.weight", "policy_fc1.bias")


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def tables_of(g):
    return {k: t(g[f"tree.{k}"]) for k in ("index", "value", "chance", "expected_value", "legal")}


def weights_of(g, prefix):
    return {k: t(g[f"{prefix}.{k}"]) for k in NET_KEYS}


def episodes_of(g):
    return {k: t(g[f"ep.{k}"]) for k in ("indices", "turns", "observations", "policy", "actions", "rewards",
                                          "values", "masks")}


def close(a, b, rtol=1e-5, atol=1e-6):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs()
    bound = atol + rtol * b.abs()
    assert bool((err <= bound).all()), f"max err {err.max().item():.3e} (worst excess {(err - bound).max().item():.3e})"


def tree_from_golden(g, device="cpu"):
    """An `environment.tree.Tree` holding the golden fixture's tables (no generation)."""
    from environment.tree import Tree

    a, c = int(g["meta"][0]), int(g["meta"][1])
    tree = Tree(max_actions=a, max_transitions=c)
    tree.index_tensor = t(g["tree.index"])
    tree.value_tensor = t(g["tree.value"])
    tree.chance_tensor = t(g["tree.chance"])
    tree.expected_value_tensor = t(g["tree.expected_value"])
    tree.legal_tensor = t(g["tree.legal"])
    tree.root_value_tensor = t(g["tree.root_value"])
    tree.solution_tensor = t(g["tree.solution"])
    tree.hash = 1234
    tree.to(torch.device(device))
    return tree


def mlp_from_golden(g, prefix, device="cpu"):
    from nn.net import MLP

    a, width = int(g["meta"][0]), int(g["meta"][3])
    net = MLP(a, width, device=torch.device(device))
    net.load_state_dict({k: v.to(device) for k, v in weights_of(g, prefix).items()})
    return net


def episodes_from_golden(g, tree, device="cpu"):
    from environment.episode import Episodes

    ep_t = episodes_of(g)
    ep = Episodes(tree, ep_t["indices"].shape[1])
    for k, v in ep_t.items():
        setattr(ep, k, v.to(device))
    ep.t_eff = int(g["ep.t_eff"])
    ep.finished = True
    ep.q_estimates = torch.zeros_like(ep.policy)
    ep.v_estimates = torch.zeros_like(ep.rewards)
    return ep


def gross_rollout_errors(ep, net, tol=2e-2):
    """Recorded observations through a plain fp32 torch forward on the episodes' device: boolean (T, B) masks of the valid
    slots whose recorded value / policy is off by more than `tol`.  The tensor-core engines differ from fp32 by < 1e-2;
    a mis-ordered tensor-memory access (an ordering race in the fused kernel) by much more.  Returns the fp32 value too."""
    import torch

    T = ep.t_eff + 1
    B = ep.indices.shape[1]
    a = ep.policy.shape[-1]
    obs = ep.observations[:T].reshape(T * B, -1)
    allow = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            hv = torch.relu(obs @ net.value_fc0.weight.T + net.value_fc0.bias)
            val = (hv @ net.value_fc1.weight.T + net.value_fc1.bias).reshape(T, B)
            hp = torch.relu(obs @ net.policy_fc0.weight.T + net.policy_fc0.bias)
            logit = (hp @ net.policy_fc1.weight.T + net.policy_fc1.bias).reshape(T, B, a)
            mask = ep.masks[:T] > 0
            e = torch.where(mask, torch.exp(logit - logit.max(-1, keepdim=True).values), torch.zeros_like(logit))
            pol = e / e.sum(-1, keepdim=True).clamp_min(1e-12)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = allow
    valid = ep.indices[:T] != 0
    bad_v = ((ep.values[:T] - val).abs() > tol) & valid
    bad_p = ((ep.policy[:T] - pol).abs().max(-1).values > tol) & valid
    return bad_v, bad_p, val
