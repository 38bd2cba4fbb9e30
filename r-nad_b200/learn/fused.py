"""
Fused learner-side net passes (csrc/learner_mlp.cu): the four `forward_batch` calls
of reference rnad.py:373-380 as one kernel, and the parameter gradients of
rnad.py:425 (`loss.backward()`) as another, for `nn.net.MLP` nets the tensor-core
engine supports (width 256, 2 <= max_actions <= 4).  The hidden activations
(T*B x 256 per trunk) stay on the SM; first layers run on tcgen05 in tf32 with fp32
accumulation, everything after them in fp32.

`RNaD.__learn` uses this engine by default where it applies; set the environment
variable RNAD_LEARNER_ENGINE=torch (or `trial.learner_engine = "torch"`) for the
reference-style fp32 batched-GEMM + autograd path.
"""

import ctypes
import os

import torch

import _b200


def supported(net) -> bool:
    from nn.net import MLP

    if type(net) is not MLP or not next(net.parameters()).is_cuda:
        return False
    return bool(_b200.lib().rnad_learner_mlp_supported(net.max_actions, net.width))


def engine_for(net, requested=None) -> str:
    requested = requested or os.environ.get("RNAD_LEARNER_ENGINE") or "auto"
    if requested == "torch":
        return "torch"
    ok = supported(net)
    if requested == "fused" and not ok:
        raise _b200.RnadError("RNAD_LEARNER_ENGINE=fused needs CUDA nn.net.MLP nets of width 256 with 2..4 actions")
    return "fused" if ok else "torch"


class FusedLearner:
    """Buffers that persist across learner steps: kernel workspace and the flat gradient the params' .grad view."""

    def __init__(self, net):
        L = _b200.lib()
        self.a, self.width = net.max_actions, net.width
        self.device = next(net.parameters()).device
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(int(L.rnad_learner_mlp_workspace_bytes(self.a, self.width)) + 256,
                                         dtype=torch.uint8, device=self.device)
            self.n_params = int(L.rnad_learner_param_count(self.a, self.width))
            self.flat_grad = torch.zeros(self.n_params, dtype=torch.float32, device=self.device)
        assert self.workspace.data_ptr() % 256 == 0
        assert self.n_params == sum(p.numel() for p in net.parameters())

    def forward(self, observations, net, net_target, net_reg, net_reg_):
        """
        observations (T,B,2,A,A).  Returns dict: logit, pi, log_pi (T,B,A), v (T,B,1) of `net`;
        v_target (T,B,1) of `net_target`; log_pi_reg, log_pi_reg_ (T,B,A).  No autograd graph.
        """
        t, b = observations.shape[:2]
        a, n, dev = self.a, t * b, self.device
        obs = observations.detach().contiguous()
        with torch.cuda.device(dev):
            out = {k: torch.empty((t, b, a), dtype=torch.float32, device=dev)
                   for k in ("logit", "pi", "log_pi", "log_pi_reg", "log_pi_reg_")}
            out["v"] = torch.empty((t, b, 1), dtype=torch.float32, device=dev)
            out["v_target"] = torch.empty((t, b, 1), dtype=torch.float32, device=dev)
            o = _b200.LearnerFwdOut(**{k: v.data_ptr() for k, v in out.items()})
            ws = [_b200.mlp_weights(x, dev) for x in (net, net_target, net_reg, net_reg_)]
            _b200.lib().rnad_learner_forward(_b200.ptr(obs, torch.float32), n, a, *[ctypes.byref(w) for w in ws],
                                             ctypes.byref(o), _b200.ptr(self.workspace), _b200.stream())
        return out

    def backward(self, observations, net, d_logit, d_v):
        """Writes the learner's parameter gradients for (d_logit (T,B,A), d_v (T,B)) into `net`'s .grad (views of flat_grad)."""
        t, b = observations.shape[:2]
        obs = observations.detach().contiguous()
        with torch.cuda.device(self.device):
            w = _b200.mlp_weights(net, self.device)
            _b200.lib().rnad_learner_backward(_b200.ptr(obs, torch.float32), t * b, self.a, ctypes.byref(w),
                                              _b200.ptr(d_logit, torch.float32), _b200.ptr(d_v, torch.float32),
                                              _b200.ptr(self.flat_grad), _b200.ptr(self.workspace), _b200.stream())
        offset = 0
        for p in net.parameters():          # registration order == state_dict order == the kernel's flat layout
            p.grad = self.flat_grad[offset: offset + p.numel()].view_as(p)
            offset += p.numel()
        return self.flat_grad
