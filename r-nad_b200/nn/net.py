"""
Actor-critic networks - API mirror of the reference `nn/net.py` (`MLP` :18-85,
`CrossConv` / `ConvResBlock` / `ConvNet` :88-269): same constructors, parameter
names (checkpoints are interchangeable), `forward`, `forward_policy`,
`forward_batch`.

`MLP` is the net the fused rollout kernel runs in-kernel (csrc/rollout_*.cu);
the methods below are its step-by-step and learner-side faces:
  * `forward` (one half-move, used by the step-by-step rollout and by users)
    samples with the project's Philox / inverse-CDF kernel instead of
    `torch.multinomial` (net.py:49);
  * `forward_batch` runs every half-move of a trajectory as ONE (T*B)-row GEMM
    per layer instead of the reference's Python loop over t (net.py:67).
`ConvNet` is outside the accelerated path and is
plain PyTorch.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F

import _b200


def _masked_policy(logits, filter_row):
    """net.py:45-46: softmax over legal actions without max-subtraction, L1-normalised with eps 1e-12."""
    exp_logits = torch.where(filter_row, torch.exp(logits), torch.zeros_like(logits))
    return exp_logits, exp_logits / exp_logits.sum(dim=-1, keepdim=True).clamp_min(1e-12)


def sample_actions(policy: torch.Tensor, u: torch.Tensor = None) -> torch.Tensor:
    """One categorical draw per row of `policy` (B, A) -> (B,) int64, on the GPU (rnad_sample_categorical)."""
    policy = policy.detach().to(torch.float32).contiguous()
    b, n = policy.shape
    with _b200.device_guard(policy):
        out = torch.empty((b,), dtype=torch.int64, device=policy.device)
        if u is not None:
            u = u.to(device=policy.device, dtype=torch.float32).contiguous()
        seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item()) if u is None else 0
        _b200.lib().rnad_sample_categorical(_b200.ptr(policy), b, n, _b200.ptr(u), seed, 0, 0, _b200.ptr(out),
                                            _b200.stream())
    return out


class MLP(nn.Module):
    def __init__(self, max_actions, width, device=torch.device("cpu:0"), dtype=torch.float):
        """Two independent one-hidden-layer trunks: value (2A^2 -> width -> 1) and policy (2A^2 -> width -> A)."""
        super().__init__()
        self.device = device
        self.value_fc0 = nn.Linear(2 * max_actions ** 2, width, device=device, dtype=dtype)
        self.value_fc1 = nn.Linear(width, 1, device=device, dtype=dtype)
        self.policy_fc0 = nn.Linear(2 * max_actions ** 2, width, device=device, dtype=dtype)
        self.policy_fc1 = nn.Linear(width, max_actions, device=device, dtype=dtype)
        self.max_actions = max_actions
        self.width = width
        self.rollout_precision = None   # None = auto (the fastest engine for the shape); "fp32" | "tf32" | "tf32x2" | "f16x2"

    def _trunks(self, flat):
        value = self.value_fc1(torch.relu(self.value_fc0(flat)))
        logits = self.policy_fc1(torch.relu(self.policy_fc0(flat)))
        return value, logits

    def forward(self, input_batch, u: torch.Tensor = None):
        """(B,2,A,A) -> logits (B,A), policy (B,A), value (B,1), sampled actions (B,) (net.py:37-51)."""
        filter_row = input_batch[:, 1, :, 0].to(torch.bool)
        flat = input_batch.reshape(-1, 2 * self.max_actions ** 2)
        value, logits = self._trunks(flat)
        _, policy = _masked_policy(logits, filter_row)
        actions = sample_actions(policy, u)
        return logits, policy, value, actions

    def forward_policy(self, input_batch: torch.Tensor) -> torch.Tensor:
        """Policy head only, with legal-action masking (net.py:53-62)."""
        filter_row = input_batch[:, 1, :, 0].to(torch.bool)
        flat = input_batch.reshape(-1, 2 * self.max_actions ** 2)
        logits = self.policy_fc1(torch.relu(self.policy_fc0(flat)))
        return _masked_policy(logits, filter_row)[1]

    def forward_batch(self, episodes):
        """[logits (T,B,A), log_policy (T,B,A), policy (T,B,A), value (T,B,1)] for a whole trajectory (net.py:64-85)."""
        obs = episodes.observations[: episodes.t_eff + 1]
        t, b = obs.shape[0], obs.shape[1]
        a = self.max_actions
        filter_row = obs[:, :, 1, :, 0].reshape(t * b, a).to(torch.bool)
        value, logits = self._trunks(obs.reshape(t * b, 2 * a * a))
        exp_logits, policy = _masked_policy(logits, filter_row)
        log_sum = torch.log(exp_logits.sum(dim=-1, keepdim=True))
        log_policy = torch.where(filter_row, logits - log_sum, torch.zeros_like(logits))
        return [logits.view(t, b, a), log_policy.view(t, b, a), policy.view(t, b, a), value.view(t, b, 1)]


class CrossConv(nn.Module):
    """Row filter + column filter spanning the whole matrix (net.py:88-143)."""

    def __init__(self, max_actions, in_channels, out_channels, device=torch.device("cpu:0"), dtype=torch.float):
        super().__init__()
        self.max_actions = max_actions
        span = 2 * max_actions - 1
        self.row_conv = nn.Conv2d(in_channels, out_channels, kernel_size=(1, span), device=device, dtype=dtype)
        self.col_conv = nn.Conv2d(in_channels, out_channels, kernel_size=(span, 1), device=device, dtype=dtype)

    def forward(self, input) -> torch.Tensor:
        p = self.max_actions - 1
        return self.row_conv(F.pad(input, (p, p, 0, 0))) + self.col_conv(F.pad(input, (0, 0, p, p)))


class ConvResBlock(nn.Module):
    """x + bn1(relu(conv1(bn0(relu(conv0(x)))))) (net.py:146-173)."""

    def __init__(self, max_actions, channels, batch_norm=False, device=torch.device("cpu:0"), dtype=torch.float):
        super().__init__()
        self.conv0 = CrossConv(max_actions, channels, channels, device=device, dtype=dtype)
        self.conv1 = CrossConv(max_actions, channels, channels, device=device, dtype=dtype)
        self.relu = torch.relu
        make_norm = (lambda: nn.BatchNorm2d(channels, device=device, dtype=dtype)) if batch_norm else nn.Identity
        self.batch_norm0 = make_norm()
        self.batch_norm1 = make_norm()

    def forward(self, input_batch) -> torch.Tensor:
        y = self.batch_norm0(self.relu(self.conv0(input_batch)))
        return input_batch + self.batch_norm1(self.relu(self.conv1(y)))


class ConvNet(nn.Module):
    """Two-headed CrossConv tower (net.py:176-269).  Stock PyTorch: not on the accelerated path."""

    def __init__(self, max_actions, channels, depth=1, batch_norm=True, device=torch.device("cpu:0"),
                 dtype=torch.float):
        super().__init__()
        self.device = device
        self.dtype = dtype
        self.max_actions = max_actions
        self.channels = channels
        self.pre = CrossConv(max_actions, in_channels=2, out_channels=channels, device=device, dtype=dtype)
        self.tower = nn.ParameterList([
            ConvResBlock(max_actions=max_actions, channels=channels, batch_norm=batch_norm, device=device, dtype=dtype)
            for _ in range(depth)
        ])
        self.policy = nn.Linear(channels * (max_actions ** 2), max_actions, device=device, dtype=dtype)
        self.value = nn.Linear(channels * (max_actions ** 2), 1, device=device, dtype=dtype)

    def _features(self, x):
        x = self.pre(x)
        for block in self.tower:
            x = block(x)
        return x.reshape(-1, self.channels * (self.max_actions ** 2))

    def _softmax_then_mask(self, logits, filter_row):
        # the ConvNet masks AFTER a full softmax (net.py:213-216), unlike the MLP
        policy = F.softmax(logits, dim=1) * filter_row
        return F.normalize(policy, dim=1, p=1)

    def forward(self, input_batch, u: torch.Tensor = None):
        x = self._features(input_batch)
        logits = self.policy(x)
        policy = self._softmax_then_mask(logits, input_batch[:, 1, :, 0])
        value = self.value(x)
        return logits, policy, value, sample_actions(policy, u)

    def forward_policy(self, input_batch) -> torch.Tensor:
        return self._softmax_then_mask(self.policy(self._features(input_batch)), input_batch[:, 1, :, 0])

    def forward_batch(self, episodes):
        obs = episodes.observations[: episodes.t_eff + 1]
        t, b = obs.shape[0], obs.shape[1]
        a = self.max_actions
        flat_obs = obs.reshape(t * b, 2, a, a)
        x = self._features(flat_obs)
        logits = self.policy(x)
        filter_row = flat_obs[:, 1, :, 0].to(torch.bool)
        exp_logits, policy = _masked_policy(logits, filter_row)
        log_sum = torch.log(exp_logits.sum(dim=-1, keepdim=True))
        log_policy = torch.where(filter_row, logits - log_sum, torch.zeros_like(logits))
        value = self.value(x)
        return [logits.view(t, b, a), log_policy.view(t, b, a), policy.view(t, b, a), value.view(t, b, 1)]
