"""
Policy post-processing, v-trace returns and the NeuRD / critic losses - API
mirror of the reference `learn/vtrace.py` (itself derived from OpenSpiel's
R-NaD), every function with the reference's name, argument order and result
shapes / dtypes.

  process_policy, v_trace        -> kernels rnad_process_policy / rnad_vtrace
                                    (csrc/vtrace.cu), one thread per game walking
                                    time backwards instead of T x ~75 ATen launches;
  learner_targets                -> the fused kernel RNaD uses: both players'
                                    v-trace, the reward transform, process_policy,
                                    the clipped NeuRD force, both loss values and
                                    their analytic gradients in ONE pass
                                    (rnad_learner_targets);
  _player_others, _policy_ratio, _has_played, get_loss_v, get_loss_nerd,
  apply_force_with_threshold, renormalize
                                 -> small differentiable torch expressions kept for
                                    API compatibility (the fused kernel supersedes
                                    them on the hot path; they are what the parity
                                    tests compare the kernel's gradients with).
"""

import ctypes
from typing import Any, Sequence, Tuple

import torch

import _b200


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def process_policy(policy: torch.Tensor, mask: torch.Tensor, n_disc, epsilon_threshold=0.03) -> torch.Tensor:
    """
    Thresholds (drops probabilities < epsilon unless all are) and discretises the
    policy to multiples of 1/n_disc (vtrace.py:24-55).  policy, mask: (T, B, A).
    No gradient flows through the result (as in the reference: ceil / int cast).
    """
    t_eff, batch_size, n_actions = policy.shape
    p, m = _f32(policy), _f32(mask)
    with _b200.device_guard(p):
        out = torch.empty_like(p)
        _b200.lib().rnad_process_policy(_b200.ptr(p), _b200.ptr(m), t_eff * batch_size, n_actions, int(n_disc),
                                        float(epsilon_threshold), _b200.ptr(out), _b200.stream())
    return out


def _player_others(player_ids: torch.Tensor, valid: torch.Tensor, player: int) -> torch.Tensor:
    """+1 where `player` moves, -1 where the opponent moves, 0 on invalid steps; shape [..., 1] (vtrace.py:70-87)."""
    res = (2 * (player_ids == player) - 1) * valid
    return torch.unsqueeze(res, dim=-1)


def _has_played(valid: torch.Tensor, player_id: torch.Tensor, player: int) -> torch.Tensor:
    """
    The reference's reverse scan (vtrace.py:141-177) never changes its carry, so
    its result is exactly `valid & (player_id == player)` as int64.
    """
    assert valid.shape == player_id.shape
    return (valid.to(torch.bool) & (player_id == player)).to(torch.long)


def _policy_ratio(pi: torch.Tensor, mu: torch.Tensor, actions_oh: torch.Tensor, valid: torch.Tensor) -> torch.Tensor:
    """pi(a)/mu(a) of the taken action, 1 on invalid steps (vtrace.py:180-204)."""
    assert pi.shape == mu.shape == actions_oh.shape

    def _select_action_prob(p):
        return torch.sum(actions_oh * p, dim=-1) * valid + (1 - valid)

    return _select_action_prob(pi) / _select_action_prob(mu)


def v_trace(
    v: torch.Tensor,
    valid: torch.Tensor,
    player_id: torch.Tensor,
    acting_policy: torch.Tensor,
    merged_policy: torch.Tensor,
    merged_log_policy: torch.Tensor,
    player_others: torch.Tensor,
    actions_oh: torch.Tensor,
    reward: torch.Tensor,
    player: int,
    # Scalars below.
    eta: float,
    lambda_: float,
    c: float,
    rho: float,
    gamma=1.0,
) -> Tuple[Any, Any, Any]:
    """
    Two-player v-trace for `player` (vtrace.py:207-352).  Returns
    v_target (T,B,1) f32, has_played (T,B) i64, learning_output (T,B,A) f32.
    """
    t_eff, batch_size, n_actions = acting_policy.shape
    args = [_f32(x) for x in (v, valid)]
    pid = player_id.detach().to(torch.long).contiguous()
    rest = [_f32(x) for x in (acting_policy, merged_policy, merged_log_policy, player_others, actions_oh, reward)]
    with _b200.device_guard(args[0]):
        dev = args[0].device
        v_target = torch.empty((t_eff, batch_size, 1), dtype=torch.float32, device=dev)
        has_played = torch.empty((t_eff, batch_size), dtype=torch.long, device=dev)
        learning_output = torch.empty((t_eff, batch_size, n_actions), dtype=torch.float32, device=dev)
        _b200.lib().rnad_vtrace(_b200.ptr(args[0]), _b200.ptr(args[1]), _b200.ptr(pid), *[_b200.ptr(x) for x in rest],
                                int(player), float(eta), float(lambda_), float(c), float(rho), float(gamma),
                                t_eff, batch_size, n_actions, _b200.ptr(v_target), _b200.ptr(has_played),
                                _b200.ptr(learning_output), _b200.stream())
    return v_target, has_played, learning_output


def apply_force_with_threshold(decision_outputs: torch.Tensor, force: torch.Tensor, threshold: float,
                               threshold_center: torch.Tensor) -> torch.Tensor:
    """Logits may only be pushed down above -threshold and up below +threshold (vtrace.py:355-367)."""
    centred = decision_outputs - threshold_center
    clipped_force = (centred > -threshold) * torch.clamp(force, max=0.0) + (centred < threshold) * torch.clamp(force, min=0.0)
    return decision_outputs * clipped_force.detach()


def renormalize(loss: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Sum of the masked loss over the number of masked steps (vtrace.py:370-374)."""
    normalization = torch.sum(mask)
    return torch.sum(loss * mask) / (normalization + (normalization == 0.0))


def get_loss_v(v_list: Sequence[torch.Tensor], v_target_list: Sequence[torch.Tensor],
               mask_list: Sequence[torch.Tensor]) -> torch.Tensor:
    """Critic loss: per player, masked mean squared error to the v-trace target (vtrace.py:377-393)."""
    total = 0
    for v_n, v_target, mask in zip(v_list, v_target_list, mask_list):
        assert v_n.shape[0] == v_target.shape[0]
        normalization = torch.sum(mask)
        total = total + torch.sum(torch.unsqueeze(mask, dim=-1) * (v_n - v_target.detach()) ** 2) / (
            normalization + (normalization == 0.0))
    return total


def get_loss_nerd(
    logit_list: Sequence[torch.Tensor],
    policy_list: Sequence[torch.Tensor],
    q_vr_list: Sequence[torch.Tensor],
    valid: torch.Tensor,
    player_ids: Sequence[torch.Tensor],
    legal_actions: torch.Tensor,
    importance_sampling_correction: Sequence[torch.Tensor],
    clip: float = 100,
    threshold: float = 2,
) -> torch.Tensor:
    """NeuRD policy loss (vtrace.py:396-431); gradient flows through the centred logits only."""
    assert isinstance(importance_sampling_correction, list)
    total = 0
    for k, (logit_pi, pi, q_vr, is_c) in enumerate(zip(logit_list, policy_list, q_vr_list,
                                                       importance_sampling_correction)):
        assert logit_pi.shape[0] == q_vr.shape[0]
        adv_pi = q_vr - torch.sum(pi * q_vr, dim=-1, keepdim=True)
        adv_pi = torch.clip(is_c * adv_pi, min=-clip, max=clip).detach()
        logits = logit_pi - torch.mean(logit_pi * legal_actions, dim=-1, keepdim=True)   # mean over all A slots
        nerd_loss = torch.sum(
            legal_actions * apply_force_with_threshold(logits, adv_pi, threshold, torch.zeros_like(logits)), dim=-1)
        total = total - renormalize(nerd_loss, valid * (player_ids == k))
    return total


def count_played(episodes) -> torch.Tensor:
    """Device int32[2]: how many valid steps each player took in the batch (the losses' normalisers N_0, N_1)."""
    # full-length tensors where a fused rollout left them (no wait for t_eff): slots past a game's end are invalid
    get = episodes.full if hasattr(episodes, "full") else (lambda key: getattr(episodes, key))
    indices = get("indices").detach().to(torch.int64).contiguous()
    turns = get("turns").detach().to(torch.int64).contiguous()
    t, b = indices.shape
    with _b200.device_guard(indices):
        counts = torch.empty(2, dtype=torch.int32, device=indices.device)
        _b200.lib().rnad_count_played(_b200.ptr(indices), _b200.ptr(turns), t, b, _b200.ptr(counts), _b200.stream())
    return counts


class LearnerTargets:
    """Result of `learner_targets`: gradients w.r.t. the learner's outputs, loss values, and optional tensors."""

    __slots__ = ("d_logit", "d_v", "losses", "counts", "pi_processed", "v_target", "has_played", "learning_output")


def learner_targets(episodes, logit, pi, log_pi, v, v_target_net, log_pi_reg, log_pi_reg_, *, alpha, eta,
                    lambda_=1.0, c=1.0, rho=1.0, gamma=1.0, epsilon_threshold=0.03, n_discrete=32, neurd_clip=10 ** 3,
                    beta=2.0, value_weight=1.0, neurd_weight=1.0, want_outputs=False, global_counts=None,
                    workspace=None) -> LearnerTargets:
    """
    Everything `RNaD.__learn` computes between the four `forward_batch` calls and
    `loss.backward()` (rnad.py:365-425) in one fused kernel, plus the analytic
    gradients of  value_weight*loss_v + neurd_weight*loss_nerd  with respect to
    the learner's `logit` (T,B,A) and `v` (T,B,1):

        torch.autograd.backward([logit, v], [out.d_logit, out.d_v.unsqueeze(-1)])

    reproduces the reference's `loss.backward()`.  `want_outputs` additionally
    materialises pi_processed / v_target / has_played / learning_output (both
    players) for inspection and parity tests.  `global_counts` (device int32[2])
    replaces the local normalisers N_p (exact data-parallel losses on ragged trees).
    """
    L = _b200.lib()
    def field(key):
        # the trajectory tensor with as many half-moves as the learner's outputs: the full-length one a fused rollout
        # wrote (available without waiting for t_eff) if `logit` was computed on it, else the stored (t_eff + 1) one
        tensor = episodes.full(key) if hasattr(episodes, "full") else getattr(episodes, key)
        return tensor if tensor.shape[0] == logit.shape[0] else getattr(episodes, key)

    t, b, a = field("policy").shape
    dev = logit.device
    res = LearnerTargets()
    with torch.cuda.device(dev):
        io = _b200.LearnerIO()
        keep = []

        def put(name, tensor, dtype=torch.float32):
            tensor = tensor.detach()
            if tensor.dtype != dtype:
                tensor = tensor.to(dtype)
            tensor = tensor.contiguous()
            keep.append(tensor)
            setattr(io, name, _b200.ptr(tensor).value)

        put("indices", field("indices"), torch.int64)
        put("turns", field("turns"), torch.int64)
        put("mu", field("policy"))
        put("actions_oh", field("actions"))
        put("rewards", field("rewards"))
        put("masks", field("masks"))
        put("logit", logit)
        put("pi", pi)
        put("log_pi", log_pi)
        put("v", v)
        put("v_target_net", v_target_net)
        put("log_pi_reg", log_pi_reg)
        put("log_pi_reg_", log_pi_reg_)
        res.d_logit = torch.empty((t, b, a), dtype=torch.float32, device=dev)
        res.d_v = torch.empty((t, b), dtype=torch.float32, device=dev)
        res.losses = torch.empty(2, dtype=torch.float32, device=dev)
        res.counts = torch.empty(2, dtype=torch.int32, device=dev)
        io.d_logit, io.d_v = res.d_logit.data_ptr(), res.d_v.data_ptr()
        io.losses, io.counts = res.losses.data_ptr(), res.counts.data_ptr()
        res.pi_processed = res.v_target = res.has_played = res.learning_output = None
        if want_outputs:
            res.pi_processed = torch.empty((t, b, a), dtype=torch.float32, device=dev)
            res.v_target = [torch.empty((t, b, 1), dtype=torch.float32, device=dev) for _ in range(2)]
            res.has_played = [torch.empty((t, b), dtype=torch.int64, device=dev) for _ in range(2)]
            res.learning_output = [torch.empty((t, b, a), dtype=torch.float32, device=dev) for _ in range(2)]
            io.pi_processed = res.pi_processed.data_ptr()
            for k in range(2):
                io.v_target[k] = res.v_target[k].data_ptr()
                io.has_played[k] = res.has_played[k].data_ptr()
                io.learning_output[k] = res.learning_output[k].data_ptr()
        if global_counts is not None:
            put("global_counts", global_counts, torch.int32)
        params = _b200.LearnerParams(float(alpha), float(eta), float(lambda_), float(c), float(rho), float(gamma),
                                     float(epsilon_threshold), int(n_discrete), float(neurd_clip), float(beta),
                                     float(value_weight), float(neurd_weight))
        if workspace is None:
            workspace = torch.zeros(int(L.rnad_learner_targets_workspace(t, b)), dtype=torch.uint8, device=dev)
        L.rnad_learner_targets(ctypes.byref(io), ctypes.byref(params), t, b, a, _b200.ptr(workspace), _b200.stream())
    return res
