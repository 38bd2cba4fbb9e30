"""
GPU parity tests of K3 (process_policy, v-trace, fused learner targets) and of the
learner step built on it, against the golden vectors recorded from the unmodified
reference and against the CPU oracle on larger seeded inputs.
Bars: process_policy and has_played bit-exact; v-trace / NeuRD floats rtol 1e-5
(north_star), gradients additionally atol 1e-7 (they are O(1/N) small).
"""
import numpy as np
import pytest
import torch

from oracle import rnad_oracle as orc
from helpers import close, episodes_from_golden, episodes_of, mlp_from_golden, t, tree_from_golden, weights_of

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cpu(x):
    return x.detach().cpu()


def dev(x):
    return t(x).to(DEV) if isinstance(x, np.ndarray) else x.to(DEV)


def test_process_policy_matches_reference(golden):
    import learn.vtrace as vtrace

    _, g = golden
    out = vtrace.process_policy(dev(g["fb.pi"]), dev(g["ep.masks"]), 32, 0.03)
    assert torch.equal(cpu(out), t(g["pi_processed"]))


@pytest.mark.parametrize("a", [2, 3, 4, 5])
def test_process_policy_random_vs_oracle(a):
    import learn.vtrace as vtrace

    gen = torch.Generator().manual_seed(a)
    T, B = 6, 4000
    mask = (torch.rand(T, B, a, generator=gen) < 0.8).float()
    mask[..., 0] = 1.0
    p = torch.rand(T, B, a, generator=gen) ** 3 * mask
    p = p / p.sum(-1, keepdim=True)
    p[0, :50] = mask[0, :50] / mask[0, :50].sum(-1, keepdim=True)        # exact ties / uniform rows
    for n_disc, eps in ((32, 0.03), (16, 0.1), (7, 0.0)):
        want = orc.process_policy(p, mask, n_disc, eps)
        got = cpu(vtrace.process_policy(p.to(DEV), mask.to(DEV), n_disc, eps))
        assert torch.equal(got, want), f"{(got != want).any(-1).sum().item()} rows differ (A={a}, n={n_disc})"


def test_v_trace_matches_reference(golden):
    import learn.vtrace as vtrace

    _, g = golden
    ep = {k: v.to(DEV) for k, v in episodes_of(g).items()}
    eta, gamma, c_bar, rho_bar, _ = (float(x) for x in g["scalars"])
    valid = (ep["indices"] != 0).float()
    for player in range(2):
        reward = ep["rewards"] if player == 0 else -ep["rewards"]
        vt, hp, lo = vtrace.v_trace(dev(g["v_target_net"]), valid, ep["turns"], ep["policy"], dev(g["pi_processed"]),
                                    dev(g["log_policy_reg"]), vtrace._player_others(ep["turns"], valid, player),
                                    ep["actions"], reward, player, eta=eta, lambda_=1.0, c=c_bar, rho=rho_bar,
                                    gamma=gamma)
        assert hp.dtype == torch.int64 and vt.shape == (*valid.shape, 1)
        assert torch.equal(cpu(hp), t(g[f"vt{player}.has_played"]))
        assert torch.equal(cpu(hp), cpu(vtrace._has_played(valid, ep["turns"], player)))
        close(cpu(vt), g[f"vt{player}.v_target"], rtol=1e-5, atol=1e-6)
        close(cpu(lo), g[f"vt{player}.learning_output"], rtol=1e-5, atol=1e-6)
        exact = torch.equal(cpu(vt), t(g[f"vt{player}.v_target"])) and torch.equal(cpu(lo), t(g[f"vt{player}.learning_output"]))
        print(f"v_trace player {player}: bit-exact={exact}")


def test_learner_targets_match_reference(golden):
    """The fused kernel against everything the reference computed from the same four nets' outputs."""
    import learn.vtrace as vtrace

    _, g = golden
    tree = tree_from_golden(g, DEV)
    ep = episodes_from_golden(g, tree, DEV)
    eta, gamma, c_bar, rho_bar, alpha = (float(x) for x in g["scalars"])
    # the regularisation nets' log-policies are not stored in the fixture: recompute them with the oracle
    _, log_pi_reg, _, _ = orc.mlp_forward_batch(weights_of(g, "reg"), t(g["ep.observations"]))
    _, log_pi_reg_, _, _ = orc.mlp_forward_batch(weights_of(g, "reg_"), t(g["ep.observations"]))
    out = vtrace.learner_targets(ep, dev(g["fb.logit"]), dev(g["fb.pi"]), dev(g["fb.log_pi"]), dev(g["fb.v"]),
                                 dev(g["v_target_net"]), log_pi_reg.to(DEV), log_pi_reg_.to(DEV), alpha=alpha, eta=eta,
                                 c=c_bar, rho=rho_bar, gamma=gamma, neurd_clip=10 ** 3, beta=2, want_outputs=True)
    assert torch.equal(cpu(out.pi_processed), t(g["pi_processed"]))
    for p in range(2):
        assert torch.equal(cpu(out.has_played[p]), t(g[f"vt{p}.has_played"]))
        close(cpu(out.v_target[p]), g[f"vt{p}.v_target"], rtol=1e-5, atol=5e-6)
        close(cpu(out.learning_output[p]), g[f"vt{p}.learning_output"], rtol=2e-5, atol=2e-5)
    counts = cpu(out.counts).tolist()
    assert counts == [int(g["vt0.has_played"].sum()), int(g["vt1.has_played"].sum())]
    losses = cpu(out.losses)
    close(losses[0], g["loss_v"], rtol=1e-5, atol=1e-6)
    close(losses[1], g["loss_nerd"], rtol=1e-5, atol=2e-6)
    close(cpu(out.d_logit), g["d_logit"], rtol=1e-5, atol=1e-7)
    close(cpu(out.d_v).unsqueeze(-1), g["d_v"], rtol=1e-5, atol=1e-7)
    assert torch.equal(cpu(vtrace.count_played(ep)), cpu(out.counts))


def test_api_losses_match_reference(golden):
    """get_loss_v / get_loss_nerd (API-compat torch expressions) and their autograd gradients on the GPU."""
    import learn.vtrace as vtrace

    _, g = golden
    ep = {k: v.to(DEV) for k, v in episodes_of(g).items()}
    valid = (ep["indices"] != 0).float()
    logit = dev(g["fb.logit"]).requires_grad_()
    v = dev(g["fb.v"]).requires_grad_()
    vts = [dev(g[f"vt{p}.v_target"]) for p in range(2)]
    hps = [dev(g[f"vt{p}.has_played"]) for p in range(2)]
    qs = [dev(g[f"vt{p}.learning_output"]) for p in range(2)]
    lv = vtrace.get_loss_v([v] * 2, vts, hps)
    ones = torch.ones_like(valid).unsqueeze(-1)
    ln = vtrace.get_loss_nerd([logit] * 2, [dev(g["pi_processed"])] * 2, qs, valid, ep["turns"], ep["masks"],
                              [ones] * 2, clip=10 ** 3, threshold=2)
    close(cpu(lv), g["loss_v"])
    close(cpu(ln), g["loss_nerd"], atol=1e-6)
    (lv + ln).backward()
    close(cpu(logit.grad), g["d_logit"], atol=1e-7)
    close(cpu(v.grad), g["d_v"], atol=1e-7)


def test_rnad_learn_gradients_match_reference(golden):
    """RNaD.__learn on the reference's episodes and nets gives the reference's parameter gradients."""
    from learn.rnad import RNaD

    name, g = golden
    tree = tree_from_golden(g, DEV)
    ep = episodes_from_golden(g, tree, DEV)
    eta, gamma, c_bar, rho_bar, alpha = (float(x) for x in g["scalars"])
    a, width = int(g["meta"][0]), int(g["meta"][3])
    trial = RNaD(tree=tree, device=torch.device(DEV), directory_name=f"pytest_{name}", eta=eta,
                 batch_size=ep.batch_size, vtrace_gamma=gamma, c_bar=c_bar, roh_bar=rho_bar,
                 net_params={"type": "MLP", "max_actions": a, "width": width})
    for attr, prefix in (("net", "learner"), ("net_target", "target"), ("net_reg", "reg"), ("net_reg_", "reg_")):
        setattr(trial, attr, mlp_from_golden(g, prefix, DEV))
    trial.net.train()
    trial.learner_engine = "torch"          # fp32 batched GEMMs + autograd: the exact-parity engine
    trial._RNaD__learn(ep, alpha)
    for k, p in trial.net.named_parameters():
        close(cpu(p.grad), g[f"rnad_grad.{k}"], rtol=2e-5, atol=2e-7)
    losses = cpu(trial.last_losses)
    close(losses[0], g["loss_v"], rtol=2e-5, atol=2e-6)
    close(losses[1], g["loss_nerd"], rtol=2e-5, atol=4e-6)


@pytest.mark.parametrize("a,T,B", [(2, 4, 3000), (3, 8, 10000), (4, 16, 2000), (5, 6, 1500), (3, 40, 600), (3, 32, 700), (2, 1, 100)])
def test_learner_targets_random_vs_oracle(a, T, B):
    """Larger synthetic trajectories with ragged validity, off-policy ratios and non-trivial scalars."""
    import learn.vtrace as vtrace

    gen = torch.Generator().manual_seed(100 + a)
    mask = (torch.rand(T, B, a, generator=gen) < 0.75).float()
    mask[..., 0] = 1.0

    def rand_policy():
        p = (torch.rand(T, B, a, generator=gen) + 0.05) * mask
        return p / p.sum(-1, keepdim=True)

    length = torch.randint(1, T + 1, (B,), generator=gen)
    indices = (torch.arange(T).view(T, 1) < length.view(1, B)).long() * torch.randint(1, 1000, (T, B), generator=gen)
    turns = (torch.arange(T) % 2).view(T, 1).expand(T, B).contiguous()
    mu, pi = rand_policy(), rand_policy()
    logit = torch.randn(T, B, a, generator=gen) * 2
    log_pi = torch.where(mask != 0, torch.log(pi.clamp_min(1e-30)), torch.zeros_like(pi))
    log_pi_reg = torch.where(mask != 0, torch.log(rand_policy().clamp_min(1e-30)), torch.zeros_like(pi))
    log_pi_reg_ = torch.where(mask != 0, torch.log(rand_policy().clamp_min(1e-30)), torch.zeros_like(pi))
    act = torch.multinomial(mu.view(-1, a), 1, generator=gen).view(T, B)
    actions = torch.nn.functional.one_hot(act, a).float()
    last = (torch.arange(T).view(T, 1) == (length - 1).view(1, B)) & (turns == 1)
    rewards = torch.where(last, torch.randint(0, 2, (T, B), generator=gen).float() * 2 - 1, torch.zeros(T, B))
    v = torch.randn(T, B, 1, generator=gen) * 0.5
    v_net = torch.randn(T, B, 1, generator=gen) * 0.5
    alpha, eta, lam, c, rho, gamma, clip, beta = 0.4, 0.2, 1.0, 0.8, 0.9, 0.95, 0.7, 1.5

    valid = (indices != 0).float()
    pi_proc = orc.process_policy(pi, mask, 32, 0.03)
    L = log_pi - (alpha * log_pi_reg + (1 - alpha) * log_pi_reg_)
    vts, hps, qs = [], [], []
    for p in range(2):
        r = rewards if p == 0 else -rewards
        vt, hp, lo = orc.v_trace(v_net, valid, turns, mu, pi_proc, L, actions, r, p, eta=eta, lambda_=lam, c=c,
                                 rho=rho, gamma=gamma)
        vts.append(vt), hps.append(hp), qs.append(lo)
    logit_g = logit.clone().requires_grad_()
    v_g = v.clone().requires_grad_()
    lv = orc.loss_v(v_g, vts, hps)
    ln = orc.loss_nerd(logit_g, pi_proc, qs, valid, turns, mask, clip, beta)
    (2.0 * lv + 0.5 * ln).backward()

    class Ep:
        pass

    ep = Ep()
    ep.indices, ep.turns, ep.policy, ep.actions = indices.to(DEV), turns.to(DEV), mu.to(DEV), actions.to(DEV)
    ep.rewards, ep.masks = rewards.to(DEV), mask.to(DEV)
    out = vtrace.learner_targets(ep, logit.to(DEV), pi.to(DEV), log_pi.to(DEV), v.to(DEV), v_net.to(DEV),
                                 log_pi_reg.to(DEV), log_pi_reg_.to(DEV), alpha=alpha, eta=eta, lambda_=lam, c=c,
                                 rho=rho, gamma=gamma, neurd_clip=clip, beta=beta, value_weight=2.0, neurd_weight=0.5,
                                 want_outputs=True)
    assert torch.equal(cpu(out.pi_processed), pi_proc)
    for p in range(2):
        assert torch.equal(cpu(out.has_played[p]), hps[p])
        close(cpu(out.v_target[p]), vts[p], rtol=1e-5, atol=1e-6)
        close(cpu(out.learning_output[p]), qs[p], rtol=1e-5, atol=2e-6)
    close(cpu(out.losses)[0], lv.detach(), rtol=1e-5, atol=1e-6)
    close(cpu(out.losses)[1], ln.detach(), rtol=1e-5, atol=1e-6)
    close(cpu(out.d_logit), logit_g.grad, rtol=1e-5, atol=1e-8)
    close(cpu(out.d_v).unsqueeze(-1), v_g.grad, rtol=1e-5, atol=1e-8)
    # without the optional outputs the gradients are the same
    lean = vtrace.learner_targets(ep, logit.to(DEV), pi.to(DEV), log_pi.to(DEV), v.to(DEV), v_net.to(DEV),
                                  log_pi_reg.to(DEV), log_pi_reg_.to(DEV), alpha=alpha, eta=eta, lambda_=lam, c=c,
                                  rho=rho, gamma=gamma, neurd_clip=clip, beta=beta, value_weight=2.0, neurd_weight=0.5)
    assert torch.equal(cpu(lean.d_logit), cpu(out.d_logit)) and torch.equal(cpu(lean.d_v), cpu(out.d_v))


def test_rnad_short_run_reduces_exploitability():
    """cfg1-sized end-to-end run on the GPU: NashConv of the target net falls well below the initial net's."""
    import random

    from environment.tree import Tree
    from learn.rnad import RNaD
    from util.metric import NashConvData

    np.random.seed(6)
    random.seed(6)
    torch.manual_seed(6)
    tree = Tree(max_actions=2, max_transitions=1, depth_bound=2)
    tree.generate()
    tree.to(torch.device(DEV))
    trial = RNaD(tree=tree, device=torch.device(DEV), directory_name=f"pytest_cfg1_{np.random.randint(1 << 30)}",
                 eta=0.2, bounds=[6], delta_m=[100], lr=1e-3, gamma_averaging=0.01, batch_size=256, logit_clip=2,
                 net_params={"type": "MLP", "max_actions": 2, "width": 256})
    trial._RNaD__initialize()
    data = NashConvData(tree)
    data.get_nashconv_from_net(tree, trial.net_target)
    before = (data.row_best[1] + data.col_best[1]).item()
    trial.run(checkpoint_mod=10 ** 6, expl_mod=1, log_mod=10 ** 6)
    assert trial.total_steps == 600 and trial.m == 6
    data = NashConvData(tree)
    data.get_nashconv_from_net(tree, trial.net_target)
    after = (data.row_best[1] + data.col_best[1]).item()
    print(f"NashConv before {before:.4f} after 600 steps {after:.4f}; history {trial.nashconv_history}")
    assert after < 0.5 * before and after < 0.2


# ----------------------------------------------------------- fused learner net passes (tcgen05, tf32)

def _four_nets(a, seed):
    from nn.net import MLP

    torch.manual_seed(seed)
    nets = []
    for _ in range(4):
        net = MLP(a, 256)
        with torch.no_grad():
            for p in net.parameters():
                p.mul_(1.5)
        nets.append(net)
    weights = [{k: v.detach().clone() for k, v in n.state_dict().items()} for n in nets]
    for n in nets:
        n.to(DEV)
        n.device = torch.device(DEV)
    return nets, weights


def _random_observations(a, T, B, seed):
    gen = torch.Generator().manual_seed(seed)
    ev = torch.rand(T, B, a, a, generator=gen) * 2 - 1
    rows = torch.randint(1, a + 1, (T, B, 1, 1), generator=gen)
    cols = torch.randint(1, a + 1, (T, B, 1, 1), generator=gen)
    legal = ((torch.arange(a).view(1, 1, a, 1) < rows) & (torch.arange(a).view(1, 1, 1, a) < cols)).float()
    return torch.stack([ev * legal, legal], dim=2).contiguous()


# (the last case: eight or nine tiles per CTA - observation and accumulator double buffers and the barrier ring wrap around)
@pytest.mark.parametrize("a,T,B", [(2, 4, 300), (3, 8, 4099), (4, 5, 2048), (3, 1, 1), (4, 12, 4100), (3, 8, 20011)])
def test_fused_learner_forward_vs_oracle(a, T, B):
    import learn.fused as fused

    nets, weights = _four_nets(a, 5 + a)
    obs = _random_observations(a, T, B, 17)
    assert fused.supported(nets[0])
    fl = fused.FusedLearner(nets[0])
    out = fl.forward(obs.to(DEV), *nets)
    ref_net = orc.mlp_forward_batch(weights[0], obs)
    ref_tgt = orc.mlp_forward_batch(weights[1], obs)
    ref_reg = orc.mlp_forward_batch(weights[2], obs)
    ref_reg_ = orc.mlp_forward_batch(weights[3], obs)
    tol = dict(rtol=0, atol=5e-3)             # fp16 operands (11-bit significand), fp32 accumulation
    close(cpu(out["logit"]), ref_net[0], **tol)
    close(cpu(out["log_pi"]), ref_net[1], **tol)
    close(cpu(out["pi"]), ref_net[2], **tol)
    close(cpu(out["v"]), ref_net[3], **tol)
    close(cpu(out["v_target"]), ref_tgt[3], **tol)
    close(cpu(out["log_pi_reg"]), ref_reg[1], **tol)
    close(cpu(out["log_pi_reg_"]), ref_reg_[1], **tol)
    mask = obs[:, :, 1, :, 0]
    assert bool((cpu(out["pi"])[mask == 0] == 0).all()) and bool((cpu(out["log_pi"])[mask == 0] == 0).all())
    close(cpu(out["pi"]).sum(-1), torch.ones(T, B), rtol=0, atol=1e-6)
    err = (cpu(out["logit"]) - ref_net[0]).abs().max().item()
    # tight check against the engine's numerics restated on the CPU (oracle.mlp_forward_tc): fp16-rounded operands in
    # both layers (pipelined kernel, both layers on the tensor core with kind::f16; A = 4 in two launches), activations
    # rounded to fp16; fp64 accumulation.  What is left is fp32 accumulation order - and, rarely, an activation on the
    # other side of an fp16 rounding boundary (see tests/test_gpu_env_rollout.py::TOL_TC).
    flat = obs.reshape(T * B, -1)
    second = "f16"
    tol_max, tol_mean = (5e-4, 5e-6)
    for key_l, key_v, wts in (("logit", "v", weights[0]), (None, "v_target", weights[1])):
        e_logit, _, e_v, _, _, _ = orc.mlp_forward_tc(wts, flat, second)
        if key_l:
            err_l = (cpu(out[key_l]).double().reshape(T * B, a) - e_logit).abs()
            assert float(err_l.max()) < tol_max and float(err_l.mean()) < tol_mean, (float(err_l.max()), float(err_l.mean()))
        err_v = (cpu(out[key_v]).double().reshape(T * B, 1) - e_v).abs()
        assert float(err_v.max()) < tol_max and float(err_v.mean()) < tol_mean, (float(err_v.max()), float(err_v.mean()))
    for key, wts in (("log_pi_reg", weights[2]), ("log_pi_reg_", weights[3])):
        e_logit, _, _, e_exp, _, _ = orc.mlp_forward_tc(wts, flat, second)
        m = flat[:, a * a: 2 * a * a: a] != 0
        e_logp = torch.where(m, e_logit - torch.log(e_exp.sum(-1, keepdim=True)), torch.zeros_like(e_logit))
        err_p = (cpu(out[key]).double().reshape(T * B, a) - e_logp).abs()
        assert float(err_p.max()) < 2 * tol_max and float(err_p.mean()) < 2 * tol_mean, (key, float(err_p.max()), float(err_p.mean()))
    err_tc = (cpu(out["logit"]).double().reshape(T * B, a) - orc.mlp_forward_tc(weights[0], flat, second)[0]).abs().max().item()
    # the warp roles of the kernel are ordered by mbarriers only: repeated launches must give the same bits (the two
    # halves of a trunk accumulate in separate columns, added in a fixed order: nothing depends on the issue order)
    first = {k: v.clone() for k, v in out.items()}
    obs_dev = obs.to(DEV)
    for it in range(25):
        again = fl.forward(obs_dev, *nets)
        for k, v in first.items():
            assert torch.equal(again[k], v), f"launch {it}: {k} differs from the first launch by {float((again[k] - v).abs().max()):.3e}"
    print(f"fused forward A={a}: max |logit error| = {err:.2e} vs fp32 net, {err_tc:.2e} vs fp16-aware oracle")


@pytest.mark.parametrize("a,T,B", [(2, 4, 300), (3, 8, 4099), (4, 5, 2048), (3, 2, 77), (3, 8, 20011), (4, 6, 20011)])
def test_fused_learner_backward_vs_autograd(a, T, B):
    import learn.fused as fused
    from nn.net import MLP

    nets, weights = _four_nets(a, 9 + a)
    obs = _random_observations(a, T, B, 23)
    gen = torch.Generator().manual_seed(3)
    d_logit = torch.randn(T, B, a, generator=gen) / (T * B)
    d_v = torch.randn(T, B, generator=gen) / (T * B)
    d_logit[:, ::3] = 0                       # invalid / opponent steps carry exact zeros
    d_v[:, ::3] = 0
    # Reference = autograd (float64) through the engine's own forward numerics (oracle.mlp_forward_tc: tf32-rounded
    # first-layer operands, straight-through), so that the relu masks agree and the comparison is tight.  Against
    # the plain fp32 net the two differ by the ~1e-3 of hidden units whose pre-activation changes sign under
    # tf32 rounding (relative gradient error ~2e-2 on this synthetic input) - checked loosely below.
    w64 = {k: v.double().requires_grad_(True) for k, v in weights[0].items()}
    bias_k = orc.tc_bias_in_k(a)
    x = orc.tf32_rna(obs.reshape(T * B, -1)).double()

    def trunk(name):
        w0 = w64[name + "_fc0.weight"]
        w0r = w0 + (orc.tf32_rna(weights[0][name + "_fc0.weight"]).double() - w0).detach()
        b0 = w64[name + "_fc0.bias"]
        b0r = b0 + (orc.tf32_rna(weights[0][name + "_fc0.bias"]).double() - b0).detach() if bias_k else b0
        h = torch.relu(x @ w0r.T + b0r)
        return h @ w64[name + "_fc1.weight"].T + w64[name + "_fc1.bias"]

    v64, logit64 = trunk("value"), trunk("policy")
    torch.autograd.backward([logit64, v64], [d_logit.reshape(T * B, a).double(), d_v.reshape(T * B, 1).double()])
    ref = MLP(a, 256)
    ref.load_state_dict(weights[0])

    class Ep:
        pass

    ep = Ep()
    ep.observations, ep.t_eff = obs, T - 1
    logit, _, _, v = ref.forward_batch(ep)
    torch.autograd.backward([logit, v], [d_logit, d_v.unsqueeze(-1)])
    fl = fused.FusedLearner(nets[0])
    flat = fl.backward(obs.to(DEV), nets[0], d_logit.to(DEV), d_v.to(DEV))
    assert flat.numel() == sum(p.numel() for p in ref.parameters())
    for (name, p_ref), p_gpu in zip(ref.named_parameters(), nets[0].parameters()):
        got, want = cpu(p_gpu.grad).double(), w64[name].grad
        rel = (got - want).norm() / want.norm().clamp_min(1e-30)
        assert rel < 1e-3, f"{name}: relative gradient error vs the tf32-aware reference {rel:.2e}"
        rel32 = (got - p_ref.grad.double()).norm() / p_ref.grad.double().norm().clamp_min(1e-30)
        assert rel32 < 6e-2, f"{name}: relative gradient error vs the fp32 net {rel32:.2e}"
        cos = torch.nn.functional.cosine_similarity(got.flatten(), p_ref.grad.double().flatten(), dim=0)
        assert cos > 0.998, f"{name}: cosine {cos:.7f}"
    # deterministic: every further call gives the same bits (fixed reduction order; and no ordering race between the
    # warp roles, which would show up as a rare difference)
    first = fl.flat_grad.clone()
    for _ in range(25):
        again = fl.backward(obs.to(DEV), nets[0], d_logit.to(DEV), d_v.to(DEV))
        assert torch.equal(again, first)


# (the last case gives every CTA eight or nine tiles: the double-buffered tile operands and every barrier parity wrap around)
@pytest.mark.parametrize("a,T,B", [(2, 4, 300), (3, 8, 4099), (3, 3, 130), (4, 5, 2048), (3, 8, 20011)])
def test_fused_learner_backward_split_vs_autograd(a, T, B):
    """
    The learner step's backward (`rnad_learner_backward_split`): one UNNORMALISED gradient per player, rows of even t
    player 0's, odd t player 1's.  max_actions <= 3 runs on the fp16-operand engine (csrc/learner_bwd_f16.cu),
    max_actions = 4 on the tf32 one; the reference is float64 autograd through the engine's own first-layer numerics
    (operands rounded like the engine's, straight-through), per player.
    """
    import learn.fused as fused
    from nn.net import MLP

    nets, weights = _four_nets(a, 5 + a)
    obs = _random_observations(a, T, B, 29)
    gen = torch.Generator().manual_seed(11)
    d_logit = torch.randn(T, B, a, generator=gen)          # unnormalised: O(1), as rnad_learner_targets leaves them
    d_v = torch.randn(T, B, generator=gen)
    d_logit[:, ::3] = 0                                    # invalid steps carry exact zeros
    d_v[:, ::3] = 0
    f16 = a <= 3
    rnd = orc.f16_rn if f16 else orc.tf32_rna
    bias_k = orc.tc_bias_in_k(a, 16 if f16 else 8)
    x = rnd(obs.reshape(T, B, -1)).double()
    fl = fused.FusedLearner(nets[0])
    got = cpu(fl.backward_split(obs.to(DEV), nets[0], d_logit.to(DEV), d_v.to(DEV))).double()
    names = [n for n, _ in MLP(a, 256).named_parameters()]
    for player in range(2):
        w64 = {k: v.double().requires_grad_(True) for k, v in weights[0].items()}

        def trunk(name):
            w0 = w64[name + "_fc0.weight"]
            w0r = w0 + (rnd(weights[0][name + "_fc0.weight"]).double() - w0).detach()
            b0 = w64[name + "_fc0.bias"]
            b0r = b0 + (rnd(weights[0][name + "_fc0.bias"]).double() - b0).detach() if bias_k else b0
            h = torch.relu(x[player::2] @ w0r.T + b0r)
            return h @ w64[name + "_fc1.weight"].T + w64[name + "_fc1.bias"]

        v64, logit64 = trunk("value"), trunk("policy")
        torch.autograd.backward([logit64, v64], [d_logit[player::2].double(), d_v[player::2].unsqueeze(-1).double()])
        offset = 0
        for name in names:
            want = w64[name].grad
            mine = got[player, offset: offset + want.numel()].view_as(want)
            offset += want.numel()
            rel = (mine - want).norm() / want.norm().clamp_min(1e-30)
            # fp16 / tf32 operands carry 11 bits: 2.4e-4 per rounded factor; the relu masks agree by construction
            assert rel < 2e-3, f"player {player} {name}: relative gradient error vs the engine-aware reference {rel:.2e}"
    first = fl.backward_split(obs.to(DEV), nets[0], d_logit.to(DEV), d_v.to(DEV)).clone()
    for _ in range(25):      # deterministic, and no ordering race between the warp roles
        assert torch.equal(fl.backward_split(obs.to(DEV), nets[0], d_logit.to(DEV), d_v.to(DEV)), first)
    # large signals saturate at fp16's largest finite value instead of turning into Inf / NaN
    if f16:
        big = fl.backward_split(obs.to(DEV), nets[0], (d_logit * 1e6).to(DEV), (d_v * 1e6).to(DEV))
        assert bool(torch.isfinite(big).all())


def test_rnad_learn_fused_engine_tracks_reference(golden):
    """Default (fused, tf32) engine on the reference's episodes: gradients within tf32 noise of the reference's."""
    from learn.rnad import RNaD
    import learn.fused as fused

    name, g = golden
    a, width = int(g["meta"][0]), int(g["meta"][3])
    tree = tree_from_golden(g, DEV)
    ep = episodes_from_golden(g, tree, DEV)
    eta, gamma, c_bar, rho_bar, alpha = (float(x) for x in g["scalars"])
    trial = RNaD(tree=tree, device=torch.device(DEV), directory_name=f"pytest_fused_{name}", eta=eta,
                 batch_size=ep.batch_size, vtrace_gamma=gamma, c_bar=c_bar, roh_bar=rho_bar,
                 net_params={"type": "MLP", "max_actions": a, "width": width})
    for attr, prefix in (("net", "learner"), ("net_target", "target"), ("net_reg", "reg"), ("net_reg_", "reg_")):
        setattr(trial, attr, mlp_from_golden(g, prefix, DEV))
    expected = "fused" if width == 256 and 2 <= a <= 4 else "torch"
    assert fused.engine_for(trial.net) == expected
    trial._RNaD__learn(ep, alpha)
    flat_got = torch.cat([cpu(p.grad).flatten() for p in trial.net.parameters()])
    flat_ref = torch.cat([t(g[f"rnad_grad.{k}"]).flatten() for k, _ in trial.net.named_parameters()])
    rel = ((flat_got - flat_ref).norm() / flat_ref.norm()).item()
    print(f"{name}: engine={expected} relative gradient error {rel:.2e}")
    assert rel < (2e-2 if expected == "fused" else 1e-4)
    losses = cpu(trial.last_losses)
    close(losses[0], g["loss_v"], rtol=1e-2, atol=1e-3)
    close(losses[1], g["loss_nerd"], rtol=1e-2, atol=2e-3)
