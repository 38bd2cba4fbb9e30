"""
Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, baskuit/R-NaD @ 0d16392) on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

What is patched, and where: nothing inside the reference.  The harness
(i) puts `tests/golden/_standin/pygambit.py` on sys.path because the real
pygambit is absent, (ii) imports the reference from a scratch copy under /tmp
(its RNaD writes `saved_runs/` next to the package; /root/reference is
read-only), (iii) replaces `torch.multinomial` *in this process* by the
project's inverse-CDF rule fed from recorded uniforms, so that sampled
trajectories are reproducible functions of (tree, weights, uniforms), and
(iv) passes b1_adam=0.0 (the reference's int default breaks Adam on torch 2.11).
"""

import os
import random
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, "_standin"))
sys.path.insert(0, REPO)
from oracle.rnad_oracle import sample_icdf  # noqa: E402  (the sampling rule; see module docstring)

SCRATCH = tempfile.mkdtemp(prefix="rnad_ref_")
REF = os.path.join(SCRATCH, "ref")
shutil.copytree("/root/reference", REF)
sys.path.insert(0, REF)

from environment.tree import Tree  # noqa: E402
from environment.episode import States, Episodes  # noqa: E402
from nn.net import MLP  # noqa: E402
import learn.vtrace as vtrace  # noqa: E402
from learn.rnad import RNaD  # noqa: E402
from util.metric import NashConvData  # noqa: E402


def seed_all(s):
    np.random.seed(s)
    random.seed(s)
    torch.manual_seed(s)


class InjectedMultinomial:
    """torch.multinomial(p, 1) -> inverse CDF of p at the next recorded uniform vector."""

    def __init__(self, uniforms):
        self.queue = list(uniforms)

    def __enter__(self):
        self.orig = torch.multinomial
        torch.multinomial = self
        return self

    def __exit__(self, *a):
        torch.multinomial = self.orig

    def __call__(self, p, num_samples=1, **kw):
        u = self.queue.pop(0)
        return sample_icdf(p, u).unsqueeze(-1)


def tree_arrays(tree):
    return {
        "index": tree.index_tensor.numpy(), "value": tree.value_tensor.numpy(),
        "chance": tree.chance_tensor.numpy(), "expected_value": tree.expected_value_tensor.numpy(),
        "legal": tree.legal_tensor.numpy(), "root_value": tree.root_value_tensor.numpy(),
        "solution": tree.solution_tensor.numpy(),
    }


def net_arrays(net, prefix):
    return {f"{prefix}.{k}": v.detach().numpy().copy() for k, v in net.state_dict().items()}


TREE_CONFIGS = {
    # name: (seed, Tree kwargs)
    "cfg1_d2a2c1": (6, dict(max_actions=2, max_transitions=1, depth_bound=2)),
    "ragged_a3c2": (0, dict(max_actions=3, max_transitions=2, transition_threshold=0.3, depth_bound=4,
                             depth_bound_lambda=lambda t: t.depth_bound - 1 - 2 * (random.random() < 0.5))),
    "regular_a3c2d3": (1, dict(max_actions=3, max_transitions=2, depth_bound=3)),
    "shrinking_a4c2": (2, dict(max_actions=4, max_transitions=2, depth_bound=3, transition_threshold=0.1,
                                row_actions_lambda=lambda t: t.row_actions - 1)),
    "c3_a3d2": (3, dict(max_actions=3, max_transitions=3, depth_bound=2, transition_threshold=0.2)),
}


def make_tree(name):
    seed, kw = TREE_CONFIGS[name]
    seed_all(seed)
    tree = Tree(**kw)
    tree.generate()
    tree.assert_index_is_tree()
    return tree


def perturbed(net, scale, seed):
    other = MLP(net.max_actions, net.width)
    other.load_state_dict(net.state_dict())
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in other.parameters():
            p.add_(scale * torch.randn(p.shape, generator=g))
    return other


def golden_for(name, batch, width, eta, gamma, c_bar, rho_bar, alpha):
    tree = make_tree(name)
    A = tree.max_actions
    out = {f"tree.{k}": v for k, v in tree_arrays(tree).items()}
    out["meta"] = np.array([A, tree.max_transitions, batch, width], dtype=np.int64)
    out["scalars"] = np.array([eta, gamma, c_bar, rho_bar, alpha], dtype=np.float64)

    seed_all(100)
    net = MLP(A, width)
    out.update(net_arrays(net, "net"))

    # ---- Episodes.generate with injected uniforms (episode.py:175-230)
    T_max = 2 * 8
    u = torch.rand(T_max, batch, 2)
    out["uniforms"] = u.numpy()
    queue = []
    for t in range(T_max):
        queue.append(u[t, :, 0])
        if t & 1:
            queue.append(u[t, :, 1])
    ep = Episodes(tree, batch)
    with InjectedMultinomial(queue):
        ep.generate(net)
    T = ep.t_eff + 1
    for key in ("indices", "turns", "observations", "policy", "actions", "rewards", "values", "masks"):
        out[f"ep.{key}"] = getattr(ep, key).numpy()
    out["ep.t_eff"] = np.array(ep.t_eff)

    # ---- States.observations / States.step replay with forced actions (episode.py:46-125)
    states = States(tree, batch)
    obs_list, idx_list, rew_list = [], [], []
    chance_queue = [u[t, :, 1] for t in range(T) if t & 1]
    with InjectedMultinomial(chance_queue):
        for t in range(T):
            idx_list.append(states.indices.clone().long())
            obs_list.append(states.observations())
            rew_list.append(states.step(ep.actions[t].argmax(-1)).clone())
    out["states.indices"] = torch.stack(idx_list).numpy()
    out["states.observations"] = torch.stack(obs_list).numpy()
    out["states.rewards"] = torch.stack(rew_list).numpy()
    out["states.final_indices"] = states.indices.long().numpy()

    # ---- MLP.forward on the t=0 and t=1 observations (net.py:37-51)
    for t in (0, 1):
        with InjectedMultinomial([u[t, :, 0]]):
            logits, policy, value, actions = net.forward(ep.observations[t])
        out[f"fwd{t}.logits"] = logits.detach().numpy()
        out[f"fwd{t}.policy"] = policy.detach().numpy()
        out[f"fwd{t}.value"] = value.detach().numpy()
        out[f"fwd{t}.actions"] = actions.numpy()

    # ---- learner maths with distinct learner / target / reg / reg_ nets (rnad.py:353-425)
    learner = perturbed(net, 0.05, 1)
    target = perturbed(net, 0.05, 2)
    reg = perturbed(net, 0.05, 3)
    reg_ = perturbed(net, 0.05, 4)
    for nm, n_ in (("learner", learner), ("target", target), ("reg", reg), ("reg_", reg_)):
        out.update(net_arrays(n_, nm))
    learner.train()
    logit, log_pi, pi, v = learner.forward_batch(ep)
    logit.retain_grad()
    v.retain_grad()
    out["fb.logit"] = logit.detach().numpy()
    out["fb.log_pi"] = log_pi.detach().numpy()
    out["fb.pi"] = pi.detach().numpy()
    out["fb.v"] = v.detach().numpy()
    pi_processed = vtrace.process_policy(pi, ep.masks, 32, 0.03)
    out["pi_processed"] = pi_processed.detach().numpy()
    player_id = ep.turns
    valid = (ep.indices != 0).to(torch.float)
    rewards = torch.stack([ep.rewards, -ep.rewards], dim=0)
    v_targets, has_played, q_list = [], [], []
    with torch.no_grad():
        _, _, _, v_tgt = target.forward_batch(ep)
        _, log_pi_reg, _, _ = reg.forward_batch(ep)
        _, log_pi_reg_, _, _ = reg_.forward_batch(ep)
        log_policy_reg = log_pi - (alpha * log_pi_reg + (1 - alpha) * log_pi_reg_)
        for player in range(2):
            vt, hp, lo = vtrace.v_trace(
                v_tgt, valid, player_id, ep.policy, pi_processed, log_policy_reg,
                vtrace._player_others(player_id, valid, player), ep.actions, rewards[player], player,
                lambda_=1.0, c=c_bar, rho=rho_bar, eta=eta, gamma=gamma)
            v_targets.append(vt)
            has_played.append(hp)
            q_list.append(lo)
            out[f"vt{player}.v_target"] = vt.numpy()
            out[f"vt{player}.has_played"] = hp.numpy()
            out[f"vt{player}.learning_output"] = lo.numpy()
    out["v_target_net"] = v_tgt.numpy()
    out["log_policy_reg"] = log_policy_reg.detach().numpy()
    loss_v = vtrace.get_loss_v([v] * 2, v_targets, has_played)
    is_vec = torch.unsqueeze(torch.ones_like(valid), dim=-1)
    loss_nerd = vtrace.get_loss_nerd([logit] * 2, [pi_processed] * 2, q_list, valid, player_id, ep.masks,
                                     [is_vec] * 2, clip=10 ** 3, threshold=2)
    (loss_v + loss_nerd).backward()
    out["loss_v"] = loss_v.detach().numpy()
    out["loss_nerd"] = loss_nerd.detach().numpy()
    out["d_logit"] = logit.grad.numpy()
    out["d_v"] = v.grad.numpy()
    for k_, p_ in learner.named_parameters():
        out[f"grad.{k_}"] = p_.grad.numpy().copy()

    # ---- the reference's own RNaD.__learn on the same episodes (pins the glue, rnad.py:353-456)
    trial = RNaD(tree=tree, device=torch.device("cpu"), directory_name=f"golden_{name}", eta=eta, batch_size=batch,
                 b1_adam=0.0, vtrace_gamma=gamma, c_bar=c_bar, roh_bar=rho_bar,
                 net_params={"type": "MLP", "max_actions": A, "width": width})
    trial._RNaD__initialize()
    for attr, src in (("net", learner), ("net_target", target), ("net_reg", reg), ("net_reg_", reg_)):
        getattr(trial, attr).load_state_dict(src.state_dict())
    trial.net.zero_grad()
    trial._RNaD__learn(ep, alpha)
    for k_, p_ in trial.net.named_parameters():
        out[f"rnad_grad.{k_}"] = p_.grad.numpy().copy()

    # ---- NashConv of the learner and of the tree's own solution (metric.py:51-175)
    data = NashConvData(tree)
    data.get_nashconv_from_net(tree, learner)
    out["nashconv.net"] = np.array((data.row_best[1] + data.col_best[1]).item())
    out["nashconv.joint_policy"] = data.joint_policy.numpy()
    out["nashconv.row_best"] = data.row_best.numpy()
    out["nashconv.col_best"] = data.col_best.numpy()
    out["nashconv.depth"] = data.depth.numpy()
    out["nashconv.reach"] = data.reach_probability.numpy()
    data = NashConvData(tree)
    data.joint_policy = tree.solution_tensor.clone()
    data.get_nashconv(tree, tree.solution_tensor)
    out["nashconv.solution"] = np.array((data.row_best[1] + data.col_best[1]).item())

    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: S={tree.index_tensor.shape[0]} T={T} B={batch} -> {os.path.getsize(path) / 1024:.0f} KiB, "
          f"valid={valid.mean():.2f} loss_v={loss_v.item():.4f} loss_nerd={loss_nerd.item():.4f}")


if __name__ == "__main__":
    #            name              B   W    eta  gamma c    rho  alpha
    golden_for("cfg1_d2a2c1",      48, 32,  0.2, 1.0, 1.0, 1.0, 0.3)
    golden_for("ragged_a3c2",      96, 64,  0.2, 0.9, 0.8, 0.9, 0.6)
    golden_for("regular_a3c2d3",   64, 256, 0.5, 1.0, 1.0, 1.0, 1.0)
    golden_for("shrinking_a4c2",   64, 48,  1.0, 0.95, 1.0, 1.0, 0.0)
    golden_for("c3_a3d2",          64, 32,  0.2, 1.0, 0.7, 1.2, 0.5)
    shutil.rmtree(SCRATCH, ignore_errors=True)
