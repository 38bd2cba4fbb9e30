"""
ctypes binding of librnad_b200.so (C ABI: include/rnad_b200.h) - the only door
between the Python API mirror (environment/, nn/, learn/) and the sm_100a
kernels.  There is no CPU or PyTorch fallback behind it: a missing library or a
non-CUDA tensor raises.

torch supplies device memory (`tensor.data_ptr()`) and the stream
(`torch.cuda.current_stream()`); nothing else of torch crosses the boundary.
"""

import ctypes
import os
from ctypes import POINTER, Structure, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.realpath(__file__))
LIB_PATH = os.environ.get("RNAD_B200_LIB") or os.path.join(_HERE, "lib", "librnad_b200.so")   # override: development builds

PREC_FP32 = 0
PREC_TF32 = 1
PREC_TF32X2 = 2
PREC_F16X2 = 3
PRECISIONS = {"fp32": PREC_FP32, "tf32": PREC_TF32, "tf32x2": PREC_TF32X2, "f16x2": PREC_F16X2}

EXPORTS = (
    "rnad_last_error", "rnad_version", "rnad_device_sm_count", "rnad_packed_strides", "rnad_tree_pack",
    "rnad_observe", "rnad_step", "rnad_sample_categorical", "rnad_rollout", "rnad_rollout_workspace_bytes", "rnad_rollout_tc_supported",
    "rnad_rollout_tc2_supported",
    "rnad_process_policy", "rnad_vtrace", "rnad_learner_targets_workspace", "rnad_count_played",
    "rnad_learner_targets", "rnad_learner_mlp_supported", "rnad_learner_mlp_workspace_bytes",
    "rnad_learner_param_count", "rnad_learner_forward", "rnad_learner_backward", "rnad_learner_backward_split",
    "rnad_learner_pack", "rnad_learner_forward_prepacked", "rnad_learner_backward_split_prepacked",
    "rnad_step_control", "rnad_step_advance", "rnad_step_advance_fetch", "rnad_learner_tail", "rnad_xchg_bytes", "rnad_xchg_create", "rnad_xchg_open", "rnad_xchg_close",
    "rnad_xchg_destroy",
)
MAX_PEERS = 16
IPC_HANDLE_BYTES = 64


class RnadError(RuntimeError):
    pass


class MlpWeights(Structure):
    _fields_ = [(n, c_void_p) for n in ("value_fc0_w", "value_fc0_b", "value_fc1_w", "value_fc1_b",
                                        "policy_fc0_w", "policy_fc0_b", "policy_fc1_w", "policy_fc1_b")] + [("width", c_int)]


class Trajectory(Structure):
    _fields_ = [(n, c_void_p) for n in ("indices", "turns", "observations", "policy", "actions", "rewards",
                                        "values", "masks", "logits", "returns")]


class LearnerIO(Structure):
    _fields_ = (
        [(n, c_void_p) for n in ("indices", "turns", "mu", "actions_oh", "rewards", "masks", "logit", "pi", "log_pi",
                                 "v", "v_target_net", "log_pi_reg", "log_pi_reg_", "d_logit", "d_v", "pi_processed")]
        + [("v_target", c_void_p * 2), ("has_played", c_void_p * 2), ("learning_output", c_void_p * 2)]
        + [("losses", c_void_p), ("counts", c_void_p), ("global_counts", c_void_p), ("unnormalised", c_int),
           ("loss_sums", c_void_p)]
    )


class LearnerFwdOut(Structure):
    _fields_ = [(n, c_void_p) for n in ("logit", "pi", "log_pi", "v", "v_target", "log_pi_reg", "log_pi_reg_")]


class LearnerParams(Structure):
    _fields_ = [("alpha", c_float), ("eta", c_float), ("lambda_", c_float), ("c", c_float), ("rho", c_float),
                ("gamma", c_float), ("eps_threshold", c_float), ("n_disc", c_int), ("neurd_clip", c_float),
                ("beta", c_float), ("value_weight", c_float), ("neurd_weight", c_float), ("alpha_dev", c_void_p)]


class StepCtrl(Structure):
    """rnad_step_ctrl (device memory; mirrored here for offsets and for reading it back)."""
    _fields_ = [("seed", c_uint64), ("alpha", c_float), ("seq", ctypes.c_uint32), ("adam_step", c_float),
                ("error", ctypes.c_uint32), ("seed_state", ctypes.c_uint32 * 2)]


class TailArgs(Structure):
    _fields_ = ([("n_params", c_int)]
                + [(n, c_void_p) for n in ("player_grads", "stats", "loss_sums", "params", "target_params", "exp_avg",
                                           "exp_avg_sq", "flat_grad", "losses", "ctrl")]
                + [(n, c_float) for n in ("lr", "beta1", "beta2", "eps", "grad_clip", "gamma_averaging",
                                          "one_minus_gamma_averaging")]
                + [("world", c_int), ("rank", c_int), ("xchg", c_void_p * MAX_PEERS), ("losses_host", c_void_p)])


_lib = None


def lib():
    """The loaded shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RnadError(
            f"{LIB_PATH} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C r-nad_b200/csrc`).  There is no fallback path.")
    L = ctypes.CDLL(LIB_PATH)
    L.rnad_last_error.restype = ctypes.c_char_p
    L.rnad_last_error.argtypes = []
    L.rnad_version.restype = c_int
    L.rnad_device_sm_count.restype = c_int
    L.rnad_packed_strides.argtypes = [c_int, c_int, POINTER(c_int), POINTER(c_int)]
    L.rnad_tree_pack.argtypes = [c_void_p] * 5 + [c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    L.rnad_observe.argtypes = [c_void_p, c_int, c_int64, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]
    L.rnad_step.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_int, c_int64,
                            c_int64, c_void_p, c_void_p, c_void_p]
    L.rnad_sample_categorical.argtypes = [c_void_p, c_int64, c_int, c_void_p, c_uint64, c_int, c_int64, c_void_p,
                                          c_void_p]
    L.rnad_rollout.argtypes = [c_void_p, c_void_p, c_int, c_int, POINTER(MlpWeights), c_int64, c_int, c_uint64,
                               c_void_p, c_int64, c_void_p, c_int, POINTER(Trajectory), c_void_p, c_void_p, c_void_p]
    L.rnad_rollout_workspace_bytes.restype = c_int64
    L.rnad_rollout_workspace_bytes.argtypes = [c_int, c_int, c_int]
    L.rnad_rollout_tc_supported.argtypes = [c_int, c_int]
    L.rnad_rollout_tc2_supported.argtypes = [c_int, c_int, c_int]
    L.rnad_process_policy.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_void_p]
    L.rnad_vtrace.argtypes = [c_void_p] * 9 + [c_int] + [c_float] * 5 + [c_int, c_int64, c_int] + [c_void_p] * 4
    L.rnad_learner_targets_workspace.restype = c_int64
    L.rnad_learner_targets_workspace.argtypes = [c_int, c_int64]
    L.rnad_count_played.argtypes = [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]
    L.rnad_learner_targets.argtypes = [POINTER(LearnerIO), POINTER(LearnerParams), c_int, c_int64, c_int, c_void_p,
                                       c_void_p]
    L.rnad_learner_mlp_supported.argtypes = [c_int, c_int]
    L.rnad_learner_mlp_workspace_bytes.restype = c_int64
    L.rnad_learner_mlp_workspace_bytes.argtypes = [c_int, c_int]
    L.rnad_learner_param_count.argtypes = [c_int, c_int]
    L.rnad_learner_forward.argtypes = [c_void_p, c_int64, c_int] + [POINTER(MlpWeights)] * 4 + [
        POINTER(LearnerFwdOut), c_void_p, c_void_p]
    L.rnad_learner_backward.argtypes = [c_void_p, c_int64, c_int, POINTER(MlpWeights), c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p]
    L.rnad_learner_backward_split.argtypes = [c_void_p, c_int, c_int64, c_int, POINTER(MlpWeights), c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_void_p]
    L.rnad_learner_backward_split_prepacked.argtypes = L.rnad_learner_backward_split.argtypes
    L.rnad_learner_forward_prepacked.argtypes = [c_void_p, c_int64, c_int] + [POINTER(MlpWeights)] * 4 + [
        POINTER(LearnerFwdOut), c_int, c_void_p, c_void_p]
    L.rnad_learner_pack.argtypes = [c_int] + [POINTER(MlpWeights)] * 4 + [c_int, c_void_p, c_void_p]
    L.rnad_step_control.argtypes = [c_void_p, c_uint64, c_float, c_void_p]
    L.rnad_step_advance.argtypes = [c_void_p, c_void_p]
    L.rnad_step_advance_fetch.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
    L.rnad_learner_tail.argtypes = [POINTER(TailArgs), c_void_p]
    L.rnad_xchg_bytes.restype = c_int64
    L.rnad_xchg_bytes.argtypes = [c_int, c_int]
    L.rnad_xchg_create.argtypes = [c_int64, POINTER(c_void_p), ctypes.c_char_p]
    L.rnad_xchg_open.argtypes = [ctypes.c_char_p, POINTER(c_void_p)]
    L.rnad_xchg_close.argtypes = [c_void_p]
    L.rnad_xchg_destroy.argtypes = [c_void_p]
    no_errcheck = ("rnad_version", "rnad_device_sm_count", "rnad_rollout_tc_supported", "rnad_rollout_tc2_supported", "rnad_learner_mlp_supported",
                   "rnad_learner_param_count")
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is c_int and name not in no_errcheck:
            fn.errcheck = _errcheck
    _lib = L
    return L


def _errcheck(result, func, args):
    if result != 0:
        raise RnadError(f"{func.__name__} failed ({result}): {_lib.rnad_last_error().decode()}")
    return result


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RnadError("the B200 kernels need CUDA tensors (no CPU fallback); got a tensor on " + str(t.device))
    if not t.is_contiguous():
        raise RnadError("tensor handed to the kernels must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RnadError(f"expected {dtype}, got {t.dtype}")
    return c_void_p(t.data_ptr())


_LAYERS = ("value_fc0", "value_fc1", "policy_fc0", "policy_fc1")


def mlp_weights(net, device=None):
    """
    rnad_mlp_weights for an nn.net.MLP (fp32, contiguous, on one CUDA device); keeps the tensors alive on the struct.
    The struct is cached on the net and rebuilt only when a parameter's storage changed.
    """
    tensors = []
    for layer in _LAYERS:
        lin = getattr(net, layer)
        tensors.append(lin.weight)
        tensors.append(lin.bias)
    key = tuple(t.data_ptr() for t in tensors) + (str(device),)
    cached = net.__dict__.get("_b200_weights")
    if cached is not None and cached[0] == key:
        return cached[1]
    w = MlpWeights()
    keep = []
    for layer in _LAYERS:
        lin = getattr(net, layer)
        for suffix, tensor in (("w", lin.weight), ("b", lin.bias)):
            tensor = tensor.detach()
            if tensor.dtype != torch.float32 or not tensor.is_cuda or (device is not None and tensor.device != device):
                raise RnadError(f"{layer}: the kernels need fp32 weights on {device or 'a CUDA device'}")
            if not tensor.is_contiguous():
                raise RnadError(f"{layer}: the kernels need contiguous weights")
            keep.append(tensor)
            setattr(w, f"{layer}_{suffix}", tensor.data_ptr())
    w.width = net.width
    w._keep = keep
    net.__dict__["_b200_weights"] = (key, w)
    return w


def device_guard(t):
    return torch.cuda.device(t.device)


class PackedTree:
    """Device-resident packed node tables of one Tree (built by rnad_tree_pack)."""

    def __init__(self, tree, key=None):
        L = lib()
        self.key = key
        self.A = int(tree.max_actions)
        self.C = int(tree.max_transitions)
        idx = tree.index_tensor
        if not idx.is_cuda:
            raise RnadError("Tree.packed(): the tree lives on " + str(idx.device) + "; move it to a CUDA device first")
        self.device = idx.device
        self.S = int(idx.shape[0])
        evs, trs = c_int(), c_int()
        L.rnad_packed_strides(self.A, self.C, ctypes.byref(evs), ctypes.byref(trs))
        self.ev_stride, self.tr_stride = evs.value, trs.value
        with device_guard(idx):
            self.ev_tab = torch.empty((self.S, self.ev_stride), dtype=torch.int32, device=self.device)
            self.tr_tab = torch.empty((self.S, self.A * self.A, self.tr_stride), dtype=torch.int32, device=self.device)
            bad = torch.zeros(1, dtype=torch.int32, device=self.device)
            L.rnad_tree_pack(ptr(idx.contiguous(), torch.int64), ptr(tree.value_tensor.contiguous(), torch.float32),
                             ptr(tree.chance_tensor.contiguous(), torch.float32),
                             ptr(tree.expected_value_tensor.contiguous(), torch.float32),
                             ptr(tree.legal_tensor.contiguous(), torch.float32), self.S, self.C, self.A,
                             ptr(self.ev_tab), ptr(self.tr_tab), ptr(bad), stream())
            flag = int(bad.item())
        if flag == 1:
            raise RnadError("tree has a legal mask that is not a prefix rectangle (reference tree.py:133 always builds one)")
        if flag == 2:
            raise RnadError("tree has a child index outside [0, S)")
        self.max_half_moves = 2 * (getattr(tree, "_depth_hint", None) or _tree_depth(idx))

    def nbytes(self):
        return self.ev_tab.numel() * 4 + self.tr_tab.numel() * 4


def _tree_depth(index_tensor):
    """Number of matrix-game levels below the root (longest root-to-terminal path)."""
    frontier = torch.ones(1, dtype=torch.int64, device=index_tensor.device)
    depth = 0
    while frontier.numel() > 0:
        depth += 1
        children = index_tensor[frontier].reshape(-1)
        frontier = children[children != 0]
    return depth
