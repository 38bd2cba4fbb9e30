"""
ORACLE - TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement (torch-CPU / numpy, fp32) of the algorithms on the R-NaD
self-play hot path of baskuit/R-NaD, one function per reference function, each
citing the reference file:line it follows.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file, and only as the checker or as the timed CPU
baseline - never as a fallback for the CUDA path.

Pinning: the reference ships no golden vectors or known-answer tests for this
path (its only test, tests/test_nashconv.py, is vacuous - SURVEY.md section 4).
The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, run in
the build container from /root/reference by `tests/golden/make_golden.py`
(seeded; third-party pygambit replaced by `tests/golden/_standin/pygambit.py`)
and committed as `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks
every function below against them.

Two things the reference cannot pin because it never seeds and samples with
`torch.multinomial` (net.py:49, episode.py:118): the random stream and the
sampling rule.  They are defined HERE and implemented identically by the CUDA
kernels: Philox4x32-10 counters -> 24-bit uniforms, inverse-CDF selection with
sequential fp32 accumulation (`philox_uniforms`, `sample_icdf`).
"""

import numpy as np
import torch

# --------------------------------------------------------------------------
# RNG + sampling rule (defined by this project; the kernels mirror it bit for bit)
# --------------------------------------------------------------------------

_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11), vectorised over numpy uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * _PHILOX_M0
            p1 = c2.astype(np.uint64) * _PHILOX_M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + _PHILOX_W0)
            k1 = np.uint32(k1 + _PHILOX_W1)
    return c0, c1, c2, c3


def philox_uniforms(seed: int, t: int, game_ids: np.ndarray):
    """
    The two uniforms game `g` consumes at half-move `t` of a rollout started
    with `seed`: counter = (g_lo, g_hi, t, 0), key = (seed_lo, seed_hi).
    Word 0 -> action draw, word 1 -> chance draw; u = (x >> 8) * 2^-24 in [0, 1).
    """
    g = np.asarray(game_ids, dtype=np.uint64)
    x0, x1, _, _ = philox4x32_10(
        (g & np.uint64(0xFFFFFFFF)).astype(np.uint32),
        (g >> np.uint64(32)).astype(np.uint32),
        np.full(g.shape, t, dtype=np.uint32),
        np.zeros(g.shape, dtype=np.uint32),
        seed & 0xFFFFFFFF,
        (seed >> 32) & 0xFFFFFFFF,
    )
    scale = np.float32(2.0 ** -24)
    return (x0 >> np.uint32(8)).astype(np.float32) * scale, (x1 >> np.uint32(8)).astype(np.float32) * scale


def sample_icdf(p: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """
    Categorical draw replacing `torch.multinomial(p, 1)` (net.py:49,
    episode.py:118): the first k with u < p[0] + ... + p[k] (fp32, accumulated
    left to right); if rounding leaves u >= the total, the last k with p[k] > 0.
    p: (B, N) f32, u: (B,) f32 in [0, 1).  Returns (B,) int64.
    """
    B, N = p.shape
    acc = torch.zeros(B, dtype=torch.float32)
    choice = torch.zeros(B, dtype=torch.int64)
    done = torch.zeros(B, dtype=torch.bool)
    for k in range(N):
        acc = acc + p[:, k]
        positive = p[:, k] > 0
        choice = torch.where(~done & positive, torch.full_like(choice, k), choice)
        done = done | (positive & (u < acc))
    return choice


# --------------------------------------------------------------------------
# K1: batched environment (reference environment/episode.py:18-125)
# --------------------------------------------------------------------------

def observe(expected_value, legal, idx, turn: int):
    """
    `States.observations` (episode.py:46-68) for a batch whose mover is `turn`
    (all games share the mover, episode.py:96-98).
    expected_value, legal: (S,1,A,A) f32; idx: (B,) int.  Returns (B,2,A,A) f32:
    channel 0 the mover's expected-payoff matrix (row: ev, col: (-ev)^T),
    channel 1 the legal mask (col: transposed).
    """
    idx = idx.long()
    ev = expected_value[idx, 0]
    lg = legal[idx, 0]
    if turn == 0:
        return torch.stack([ev, lg], dim=1)
    return torch.stack([(-ev).transpose(1, 2), lg.transpose(1, 2)], dim=1).contiguous()


def mover_mask(obs):
    """Legal-action mask of the mover, `obs[:, 1, :, 0]` (net.py:38, episode.py:208)."""
    return obs[:, 1, :, 0]


def step(index, value, chance, idx, row_actions, col_actions, u_chance):
    """
    The transition half of `States.step` (episode.py:102-121): chance
    distribution chance[s,:,r,c], chance action k by `sample_icdf` (in place of
    torch.multinomial), s' = index[s,k,r,c], reward = value[s,k,r,c]*(s'==0).
    Returns (new_idx int64 (B,), reward f32 (B,), chance_action int64 (B,)).
    """
    idx = idx.long()
    b = torch.arange(idx.shape[0])
    probs = chance[idx][b, :, row_actions, col_actions]        # (B, C)
    k = sample_icdf(probs, u_chance)
    new_idx = index[idx][b, k, row_actions, col_actions]
    reward = value[idx][b, k, row_actions, col_actions] * (new_idx == 0)
    return new_idx, reward, k


# --------------------------------------------------------------------------
# K2: policy/value net (reference nn/net.py:18-85) and the rollout loop
# --------------------------------------------------------------------------

def mlp_forward(w, obs_flat):
    """
    `MLP.forward` without the sampling (net.py:37-47).  w: dict of the
    nn.Linear parameters value_fc0/1, policy_fc0/1 (.weight/.bias); obs_flat:
    (N, 2A^2) f32.  Returns logits (N,A), policy (N,A), value (N,1),
    exp_logits (N,A).  The mover mask is obs channel 1, column 0.
    """
    A = w["policy_fc1.weight"].shape[0]
    mask = obs_flat[:, A * A: 2 * A * A: A] != 0
    hv = torch.relu(obs_flat @ w["value_fc0.weight"].T + w["value_fc0.bias"])
    value = hv @ w["value_fc1.weight"].T + w["value_fc1.bias"]
    hp = torch.relu(obs_flat @ w["policy_fc0.weight"].T + w["policy_fc0.bias"])
    logits = hp @ w["policy_fc1.weight"].T + w["policy_fc1.bias"]
    e = torch.where(mask, torch.exp(logits), torch.zeros_like(logits))
    policy = e / torch.clamp_min(e.sum(-1, keepdim=True), 1e-12)   # F.normalize(p=1), eps 1e-12
    return logits, policy, value, e


def tf32_rna(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> tf32 (10-bit mantissa) as PTX `cvt.rna.tf32.f32`: nearest, ties away from zero."""
    bits = x.detach().to(torch.float32).contiguous().view(torch.int32)
    out = ((bits + 0x1000) & ~0x1FFF).view(torch.float32)
    return torch.where(torch.isfinite(x), out, x)


def tf32_trunc(x: torch.Tensor) -> torch.Tensor:
    """What the tensor core does to an fp32 operand it is handed as tf32: the low 13 mantissa bits are ignored."""
    bits = x.detach().to(torch.float32).contiguous().view(torch.int32)
    return (bits & ~0x1FFF).view(torch.float32)


def f16_rn(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> fp16 -> fp32 as PTX `cvt.rn.f16.f32`: nearest, ties to even (11-bit significand, like tf32)."""
    return x.detach().to(torch.float32).to(torch.float16).to(torch.float32)


def tc_bias_in_k(A: int, k_step: int = 8) -> bool:
    """
    Whether the tensor-core engines carry the first-layer bias as a constant-1 input column (then it is
    rounded like a weight) or add it in fp32 after the MMA: in K when 2A^2 is not a multiple of the
    MMA's K (8 for kind::tf32, 16 for kind::f16), i.e. when K has padding to spare.
    """
    return (2 * A * A) % k_step != 0


def mlp_forward_tc(w, obs_flat, second_layer="fp32", dtype=torch.float64):
    """
    Emulation of the NUMERICS of the tensor-core engines for `MLP.forward` (net.py:37-47), evaluated in
    float64 so that only the engines' accumulation order is left as a difference:
      first layers   x, W (and the bias when it rides in K) rounded to tf32 (cvt.rna), exact products,
                     wide accumulation;
      second layers  second_layer="fp32" (precision "tf32" rollout engine, learner kernels): fp32 operands;
                     second_layer="tf32" (precision "tf32x2" rollout engine): relu(h) in fp32 truncated to
                     tf32 by the tensor core, W rounded to tf32 (cvt.rna), bias added in fp32.
      second_layer="f16" (precision "f16x2", kind::f16): x, W, the bias when it rides in K (2A^2 not a multiple of
                     16), relu(h) and the second-layer W all rounded to fp16 (cvt.rn), fp32 accumulation, the second
                     bias added in fp32.
    Returns logits, policy, value, exp_logits, hidden_value, hidden_policy (pre-activation) in `dtype`.
    """
    A = w["policy_fc1.weight"].shape[0]
    f16 = second_layer == "f16"
    rnd = f16_rn if f16 else tf32_rna
    bias_k = tc_bias_in_k(A, 16 if f16 else 8)
    x = rnd(obs_flat).to(dtype)
    mask = obs_flat[:, A * A: 2 * A * A: A] != 0

    def trunk(name):
        w0 = rnd(w[name + "_fc0.weight"]).to(dtype)
        b0 = (rnd(w[name + "_fc0.bias"]) if bias_k else w[name + "_fc0.bias"]).to(dtype)
        pre = x @ w0.T + b0
        h = torch.relu(pre)
        if f16:
            h = f16_rn(h.to(torch.float32)).to(dtype)
            w1 = f16_rn(w[name + "_fc1.weight"]).to(dtype)
        elif second_layer == "tf32":
            h = tf32_trunc(h.to(torch.float32)).to(dtype)
            w1 = tf32_rna(w[name + "_fc1.weight"]).to(dtype)
        else:
            w1 = w[name + "_fc1.weight"].to(dtype)
        return pre, h @ w1.T + w[name + "_fc1.bias"].to(dtype)

    pre_v, value = trunk("value")
    pre_p, logits = trunk("policy")
    e = torch.where(mask, torch.exp(logits), torch.zeros_like(logits))
    policy = e / torch.clamp_min(e.sum(-1, keepdim=True), 1e-12)
    return logits, policy, value, e, pre_v, pre_p


def mlp_forward_batch(w, observations):
    """
    `MLP.forward_batch` (net.py:64-85) on (T,B,2,A,A) observations, all T at
    once.  Returns logits (T,B,A), log_policy (T,B,A), policy (T,B,A), value (T,B,1).
    """
    T, B = observations.shape[:2]
    A = observations.shape[-1]
    flat = observations.reshape(T * B, 2 * A * A)
    logits, policy, value, e = mlp_forward(w, flat)
    mask = flat[:, A * A: 2 * A * A: A] != 0
    log_policy = torch.where(mask, logits - torch.log(e.sum(-1, keepdim=True)), torch.zeros_like(logits))
    return (logits.view(T, B, A), log_policy.view(T, B, A), policy.view(T, B, A), value.view(T, B, 1))


def rollout(tables, w, batch_size, max_half_moves, seed=None, uniforms=None, game_offset=0):
    """
    `Episodes.generate` (episode.py:175-230): alternate row / column half-moves
    from the root until every game sits on the absorbing node (or
    `max_half_moves`).  tables: dict index/value/chance/expected_value/legal in
    the reference layout.  Randomness: `uniforms` (T,B,2) f32 if given, else
    Philox (`seed`, game id = game_offset + b).
    Returns a dict with the reference's (T,B,...) trajectory tensors.
    """
    A = tables["legal"].shape[-1]
    B = batch_size
    idx = torch.ones(B, dtype=torch.int64)
    rec = {k: [] for k in ("indices", "turns", "observations", "policy", "actions", "rewards", "values", "masks")}
    row_a = None
    games = np.arange(game_offset, game_offset + B)
    t = 0
    while t < max_half_moves and bool((idx != 0).any()):
        turn = t & 1
        if uniforms is not None:
            u_act, u_ch = uniforms[t, :, 0], uniforms[t, :, 1]
        else:
            ua, uc = philox_uniforms(seed, t, games)
            u_act, u_ch = torch.from_numpy(ua), torch.from_numpy(uc)
        obs = observe(tables["expected_value"], tables["legal"], idx, turn)
        logits, policy, value, _ = mlp_forward(w, obs.reshape(B, -1))
        act = sample_icdf(policy, u_act)
        rec["indices"].append(idx.clone())
        rec["turns"].append(torch.full((B,), turn, dtype=torch.int64))
        rec["observations"].append(obs)
        rec["policy"].append(policy)
        rec["actions"].append(torch.nn.functional.one_hot(act, A).float())
        rec["values"].append(value[:, 0])
        rec["masks"].append(mover_mask(obs).clone())
        if turn == 0:
            row_a = act
            rec["rewards"].append(torch.zeros(B))
        else:
            idx, reward, _ = step(tables["index"], tables["value"], tables["chance"], idx, row_a, act, u_ch)
            rec["rewards"].append(reward)
        t += 1
    out = {k: torch.stack(v, 0) for k, v in rec.items()}
    out["t_eff"] = t - 1
    return out


# --------------------------------------------------------------------------
# K3: policy post-processing, v-trace, NeuRD / critic losses
#     (reference learn/vtrace.py, learn/rnad.py:353-425)
# --------------------------------------------------------------------------

def process_policy(policy, mask, n_disc=32, epsilon_threshold=0.03):
    """
    `vtrace.process_policy` (vtrace.py:24-55): drop probabilities below the
    threshold (unless all are), renormalise, then hand out n_disc blocks of
    mass 1/n_disc in descending-probability order, ceil(n_disc*p) each while
    blocks last.  Ties are broken towards the lower action id (the reference's
    argsort is unstable; its tests never hit exact ties).
    """
    shape = policy.shape
    A = shape[-1]
    p = policy.reshape(-1, A)
    m = mask.reshape(-1, A)
    keep = m * ((p >= epsilon_threshold) + (p.max(-1, keepdim=True).values < epsilon_threshold))
    q = keep * p / (keep * p).sum(-1, keepdim=True)
    blocks = torch.ceil(n_disc * q)
    order = torch.sort(q, dim=-1, descending=True, stable=True).indices
    left = torch.full((p.shape[0],), float(n_disc))
    out = torch.zeros_like(q)
    rows = torch.arange(p.shape[0])
    for i in range(A):
        a = order[:, i]
        x = torch.minimum(left, blocks[rows, a])
        left = left - x
        out[rows, a] += x
    return (out / n_disc).view(shape)


def v_trace(v, valid, player_id, acting_policy, merged_policy, merged_log_policy, actions_oh, reward, player,
            eta, lambda_=1.0, c=1.0, rho=1.0, gamma=1.0):
    """
    `vtrace.v_trace` for one player (vtrace.py:207-352, helpers 70-87, 141-204),
    written as the explicit reverse recurrence:
    carry (R, Ru, nv, nvt, IS) = (0,0,0,0,1) and for t = T-1 .. 0
        Ru2 = r + gamma*Ru + ent ;  dR = r + gamma*R
        vt  = v + min(cs*IS, rho)*(Ru2 + gamma*nv - v) + lambda*min(cs*IS, c)*gamma*(nvt - nv)
        lo  = v + elp + a_oh*inv_mu*(dR + gamma*IS*nvt - v)
        own  : emit (vt, lo), carry <- (0, 0, v, vt, 1)
        opp  : emit 0,        carry <- (ent + cs*dR, Ru2, gamma*nv, gamma*nvt, cs*IS)
        else : emit 0,        carry <- (0,0,0,0,1)
    Shapes: v (T,B,1); valid (T,B) f32; player_id (T,B) i64; policies/log
    policy/actions_oh (T,B,A); reward (T,B).  Returns v_target (T,B,1),
    has_played (T,B) i64, learning_output (T,B,A).
    """
    T, B, A = acting_policy.shape
    own = (player_id == player)
    po = (2.0 * own.float() - 1.0) * valid                             # _player_others, vtrace.py:70-87

    def sel(pi):                                                       # _policy_ratio, vtrace.py:180-204
        return (actions_oh * pi).sum(-1) * valid + (1 - valid)

    mu_a = sel(acting_policy)
    cs = sel(merged_policy) / mu_a
    inv_mu = sel(torch.ones_like(merged_policy)) / mu_a
    ent = -eta * (merged_policy * merged_log_policy).sum(-1) * po      # vtrace.py:234-238
    elp = -eta * merged_log_policy * po.unsqueeze(-1)                  # vtrace.py:239

    R = torch.zeros(B)
    Ru = torch.zeros(B)
    nv = torch.zeros(B)
    nvt = torch.zeros(B)
    IS = torch.ones(B)
    v_target = torch.zeros(T, B)
    lo_out = torch.zeros(T, B, A)
    is_valid = valid != 0
    for t in range(T - 1, -1, -1):
        vv = v[t, :, 0]
        Ru2 = reward[t] + gamma * Ru + ent[t]
        dR = reward[t] + gamma * R
        w = cs[t] * IS
        vt = (vv + torch.clamp(w, max=rho) * (Ru2 + gamma * nv - vv)
              + lambda_ * torch.clamp(w, max=c) * gamma * (nvt - nv))
        lo = (vv.unsqueeze(-1) + elp[t]
              + actions_oh[t] * inv_mu[t].unsqueeze(-1) * (dR + gamma * IS * nvt - vv).unsqueeze(-1))
        mine = is_valid[t] & own[t]
        theirs = is_valid[t] & ~own[t]
        zero = torch.zeros(B)
        one = torch.ones(B)
        v_target[t] = torch.where(mine, vt, zero)
        lo_out[t] = torch.where(mine.unsqueeze(-1), lo, torch.zeros_like(lo))
        R_n = torch.where(mine, zero, torch.where(theirs, ent[t] + cs[t] * dR, zero))
        Ru_n = torch.where(mine, zero, torch.where(theirs, Ru2, zero))
        nv_n = torch.where(mine, vv, torch.where(theirs, gamma * nv, zero))
        nvt_n = torch.where(mine, vt, torch.where(theirs, gamma * nvt, zero))
        IS_n = torch.where(mine, one, torch.where(theirs, w, one))
        R, Ru, nv, nvt, IS = R_n, Ru_n, nv_n, nvt_n, IS_n
    has_played = (is_valid & own).long()                               # _has_played, vtrace.py:141-177
    return v_target.unsqueeze(-1), has_played, lo_out


def loss_v(v, v_targets, has_played):
    """`vtrace.get_loss_v` (vtrace.py:377-393) for the two players; v carries grad."""
    total = 0
    for vt, hp in zip(v_targets, has_played):
        hp = hp.float()
        n = hp.sum()
        total = total + (hp.unsqueeze(-1) * (v - vt) ** 2).sum() / (n + (n == 0))
    return total


def neurd_force(logits, pi_processed, q, legal, neurd_clip, beta):
    """
    Clipped NeuRD force and centred logits for one player (vtrace.py:355-367,
    415-422): adv = clip(q - sum_a pi~ q); lc = logit - mean_A(logit*legal);
    force = [lc > -beta]*min(adv,0) + [lc < beta]*max(adv,0).
    """
    adv = q - (pi_processed * q).sum(-1, keepdim=True)
    adv = torch.clamp(adv, min=-neurd_clip, max=neurd_clip)
    lc = logits - (logits * legal).mean(-1, keepdim=True)
    force = (lc > -beta) * torch.clamp(adv, max=0.0) + (lc < beta) * torch.clamp(adv, min=0.0)
    return lc, force


def loss_nerd(logits, pi_processed, q_list, valid, player_id, legal, neurd_clip, beta):
    """`vtrace.get_loss_nerd` (vtrace.py:396-431) with importance_sampling_correction == 1."""
    total = 0
    for k, q in enumerate(q_list):
        lc, force = neurd_force(logits, pi_processed, q, legal, neurd_clip, beta)
        per_state = (legal * lc * force.detach()).sum(-1)
        m = valid * (player_id == k)
        n = m.sum()
        total = total - (per_state * m).sum() / (n + (n == 0))
    return total


def learner_targets(w_net, w_target, w_reg, w_reg_, ep, alpha, eta, n_disc=32, eps_thr=0.03, neurd_clip=1e3,
                    beta=2.0, c_bar=1.0, rho_bar=1.0, gamma=1.0):
    """
    Everything `RNaD.__learn` (rnad.py:353-425) computes up to the loss, plus
    the analytic gradients the autograd call at rnad.py:425 produces w.r.t. the
    learner net's `logit` and `v` outputs.
    ep: dict of (T,B,...) trajectory tensors (indices, turns, observations,
    policy, actions, rewards, masks).
    """
    player_id = ep["turns"]
    valid = (ep["indices"] != 0).float()
    masks = ep["masks"]
    logit, log_pi, pi, v = mlp_forward_batch(w_net, ep["observations"])
    pi_proc = process_policy(pi, masks, n_disc, eps_thr)
    _, _, _, v_tgt = mlp_forward_batch(w_target, ep["observations"])
    _, log_pi_reg, _, _ = mlp_forward_batch(w_reg, ep["observations"])
    _, log_pi_reg_, _, _ = mlp_forward_batch(w_reg_, ep["observations"])
    log_policy_reg = log_pi - (alpha * log_pi_reg + (1 - alpha) * log_pi_reg_)        # rnad.py:382
    rewards = (ep["rewards"], -ep["rewards"])                                          # rnad.py:368
    v_targets, has_played, q_list = [], [], []
    for player in range(2):
        vt, hp, lo = v_trace(v_tgt, valid, player_id, ep["policy"], pi_proc, log_policy_reg, ep["actions"],
                             rewards[player], player, eta=eta, lambda_=1.0, c=c_bar, rho=rho_bar, gamma=gamma)
        v_targets.append(vt)
        has_played.append(hp)
        q_list.append(lo)
    lv = loss_v(v, v_targets, has_played)
    ln = loss_nerd(logit, pi_proc, q_list, valid, player_id, masks, neurd_clip, beta)

    A = logit.shape[-1]
    d_v = torch.zeros_like(v)
    g = torch.zeros_like(logit)
    for k in range(2):
        hp = has_played[k].float()
        n = torch.clamp_min(hp.sum(), 1.0)
        d_v = d_v + 2.0 * hp.unsqueeze(-1) * (v - v_targets[k]) / n
        _, force = neurd_force(logit, pi_proc, q_list[k], masks, neurd_clip, beta)
        g = g - masks * force * hp.unsqueeze(-1) / n
    d_logit = g - masks * g.sum(-1, keepdim=True) / A
    return {
        "logit": logit, "log_pi": log_pi, "pi": pi, "v": v, "pi_processed": pi_proc, "v_target_net": v_tgt,
        "log_policy_reg": log_policy_reg, "v_targets": v_targets, "has_played": has_played,
        "learning_outputs": q_list, "loss_v": lv, "loss_nerd": ln, "d_v": d_v, "d_logit": d_logit,
    }
