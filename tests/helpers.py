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
