"""
Pins oracle/rnad_oracle.py against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).
Index / mask / action / reward results must be bit-exact; floats that go
through a GEMM are compared at rtol 1e-5 (torch-CPU sgemm blocking differs
between the reference's per-t calls and the oracle's batched call).
"""
import numpy as np
import torch

from oracle import rnad_oracle as orc
from helpers import close, episodes_of, t, tables_of, weights_of


def test_observe_step_replay(golden):
    _, g = golden
    tab = tables_of(g)
    idx_ref = t(g["states.indices"])
    T, B = idx_ref.shape
    actions = t(g["ep.actions"]).argmax(-1)
    u = t(g["uniforms"])
    idx = torch.ones(B, dtype=torch.int64)
    row = None
    for s in range(T):
        assert torch.equal(idx, idx_ref[s])
        obs = orc.observe(tab["expected_value"], tab["legal"], idx, s & 1)
        assert torch.equal(obs, t(g["states.observations"][s]))          # value-equal incl. -0.0 == 0.0
        assert torch.equal(orc.mover_mask(obs), t(g["ep.masks"][s]))
        if s & 1:
            idx, rew, _ = orc.step(tab["index"], tab["value"], tab["chance"], idx, row, actions[s], u[s, :, 1])
        else:
            row, rew = actions[s], torch.zeros(B)
        assert torch.equal(rew, t(g["states.rewards"][s]))
    assert torch.equal(idx, t(g["states.final_indices"]))


def test_mlp_forward(golden):
    _, g = golden
    w = weights_of(g, "net")
    for s in (0, 1):
        obs = t(g["ep.observations"][s])
        logits, policy, value, _ = orc.mlp_forward(w, obs.reshape(obs.shape[0], -1))
        close(logits, g[f"fwd{s}.logits"])
        close(policy, g[f"fwd{s}.policy"])
        close(value, g[f"fwd{s}.value"])
        act = orc.sample_icdf(t(g[f"fwd{s}.policy"]), t(g["uniforms"][s, :, 0]))
        assert torch.equal(act, t(g[f"fwd{s}.actions"]))


def test_rollout_uniform_injection(golden):
    _, g = golden
    tab, w, ep = tables_of(g), weights_of(g, "net"), episodes_of(g)
    B = ep["indices"].shape[1]
    out = orc.rollout(tab, w, B, 64, uniforms=t(g["uniforms"]))
    assert out["t_eff"] == int(g["ep.t_eff"])
    # a sampled action can only differ where the uniform sits within fp32 noise of a CDF edge
    same = torch.equal(out["actions"], ep["actions"]) and torch.equal(out["indices"], ep["indices"])
    assert same, "trajectory diverged from the reference under identical uniforms"
    assert torch.equal(out["turns"], ep["turns"])
    assert torch.equal(out["rewards"], ep["rewards"])
    assert torch.equal(out["masks"], ep["masks"])
    assert torch.equal(out["observations"], ep["observations"])
    close(out["policy"], ep["policy"])
    close(out["values"], ep["values"])


def test_forward_batch(golden):
    _, g = golden
    ep = episodes_of(g)
    logit, log_pi, pi, v = orc.mlp_forward_batch(weights_of(g, "learner"), ep["observations"])
    close(logit, g["fb.logit"])
    close(log_pi, g["fb.log_pi"])
    close(pi, g["fb.pi"])
    close(v, g["fb.v"])


def test_process_policy(golden):
    _, g = golden
    out = orc.process_policy(t(g["fb.pi"]), t(g["ep.masks"]), 32, 0.03)
    assert torch.equal(out, t(g["pi_processed"]))
    close(out.sum(-1), torch.ones(out.shape[:2]), rtol=0, atol=1e-6)


def test_v_trace_both_players(golden):
    _, g = golden
    ep = episodes_of(g)
    eta, gamma, c_bar, rho_bar, _ = g["scalars"]
    valid = (ep["indices"] != 0).float()
    for player in range(2):
        reward = ep["rewards"] if player == 0 else -ep["rewards"]
        vt, hp, lo = orc.v_trace(t(g["v_target_net"]), valid, ep["turns"], ep["policy"], t(g["pi_processed"]),
                                 t(g["log_policy_reg"]), ep["actions"], reward, player, eta=float(eta),
                                 lambda_=1.0, c=float(c_bar), rho=float(rho_bar), gamma=float(gamma))
        assert torch.equal(hp, t(g[f"vt{player}.has_played"]))
        assert torch.equal(vt, t(g[f"vt{player}.v_target"]))             # same fp32 op order -> bit-exact
        assert torch.equal(lo, t(g[f"vt{player}.learning_output"]))


def test_losses_and_gradients(golden):
    _, g = golden
    ep = episodes_of(g)
    valid = (ep["indices"] != 0).float()
    logit = t(g["fb.logit"]).requires_grad_()
    v = t(g["fb.v"]).requires_grad_()
    vts = [t(g[f"vt{p}.v_target"]) for p in range(2)]
    hps = [t(g[f"vt{p}.has_played"]) for p in range(2)]
    qs = [t(g[f"vt{p}.learning_output"]) for p in range(2)]
    lv = orc.loss_v(v, vts, hps)
    ln = orc.loss_nerd(logit, t(g["pi_processed"]), qs, valid, ep["turns"], ep["masks"], 1e3, 2.0)
    close(lv, g["loss_v"])
    close(ln, g["loss_nerd"], atol=1e-6)
    (lv + ln).backward()
    close(logit.grad, g["d_logit"], atol=1e-7)
    close(v.grad, g["d_v"], atol=1e-7)


def test_learner_targets_end_to_end(golden):
    _, g = golden
    ep = episodes_of(g)
    eta, gamma, c_bar, rho_bar, alpha = (float(x) for x in g["scalars"])
    r = orc.learner_targets(weights_of(g, "learner"), weights_of(g, "target"), weights_of(g, "reg"),
                            weights_of(g, "reg_"), ep, alpha, eta, c_bar=c_bar, rho_bar=rho_bar, gamma=gamma)
    close(r["log_policy_reg"], g["log_policy_reg"], atol=2e-6)
    close(r["v_target_net"], g["v_target_net"])
    assert torch.equal(r["pi_processed"], t(g["pi_processed"]))
    for p in range(2):
        close(r["v_targets"][p], g[f"vt{p}.v_target"], atol=5e-6)
        close(r["learning_outputs"][p], g[f"vt{p}.learning_output"], rtol=2e-5, atol=2e-5)
    close(r["loss_v"], g["loss_v"])
    close(r["loss_nerd"], g["loss_nerd"], atol=2e-6)
    close(r["d_logit"], g["d_logit"], atol=1e-7)      # analytic gradients == autograd of the reference
    close(r["d_v"], g["d_v"], atol=1e-7)


def test_reference_rnad_learn_matches_handwired_glue(golden):
    """RNaD.__learn's own parameter gradients equal those of the hand-wired call sequence."""
    _, g = golden
    for k in ("value_fc0.weight", "policy_fc0.weight", "policy_fc1.bias", "value_fc1.weight"):
        close(g[f"rnad_grad.{k}"], g[f"grad.{k}"], atol=1e-7)


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32-10, counter 0 / key 0 and the all-ones vector
    out = orc.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    out = orc.philox4x32_10(f, f, f, f, f, f)
    assert [int(x) for x in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    out = orc.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)
    assert [int(x) for x in out] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    ua, uc = orc.philox_uniforms(7, 3, np.arange(1000))
    assert ua.min() >= 0 and ua.max() < 1 and abs(ua.mean() - 0.5) < 0.05 and abs(uc.mean() - 0.5) < 0.05


def test_sample_icdf_rule():
    p = torch.tensor([[0.25, 0.0, 0.75], [0.0, 1.0, 0.0], [0.5, 0.5, 0.0]])
    assert orc.sample_icdf(p, torch.tensor([0.2499, 0.999, 0.5])).tolist() == [0, 1, 1]
    assert orc.sample_icdf(p, torch.tensor([0.25, 0.0, 0.99999994])).tolist() == [2, 1, 1]
