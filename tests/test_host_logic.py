"""
CPU tests of the host-side logic: tree generation invariants, the vectorised
NashConv against the reference's recursive one (golden vectors), the C-ABI
library's exported symbols, and loud failure of the product path without CUDA.
"""
import ctypes
import os
import random
import re

import numpy as np
import pytest
import torch

from helpers import close, mlp_from_golden, t, tree_from_golden

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def seeded_tree(seed=0, **kw):
    from environment.tree import Tree

    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)
    tree = Tree(**kw)
    tree.generate()
    return tree


def test_generated_tree_matches_reference_generation(golden):
    """Same seeds -> same tree as the reference's generator (index / chance / legal exact, values to solver rounding)."""
    import sys
    sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
    name, g = golden
    seeds = {"cfg1_d2a2c1": (6, dict(max_actions=2, max_transitions=1, depth_bound=2)),
             "regular_a3c2d3": (1, dict(max_actions=3, max_transitions=2, depth_bound=3)),
             "c3_a3d2": (3, dict(max_actions=3, max_transitions=3, depth_bound=2, transition_threshold=0.2))}
    if name not in seeds:
        pytest.skip("tree uses lambdas; covered by the ragged cases below")
    seed, kw = seeds[name]
    tree = seeded_tree(seed, **kw)
    tree.assert_index_is_tree()
    assert torch.equal(tree.index_tensor, t(g["tree.index"]))
    assert torch.equal(tree.chance_tensor, t(g["tree.chance"]))
    assert torch.equal(tree.legal_tensor, t(g["tree.legal"]))
    close(tree.value_tensor, g["tree.value"], atol=1e-6)
    close(tree.expected_value_tensor, g["tree.expected_value"], atol=1e-6)
    close(tree.root_value_tensor, g["tree.root_value"], atol=1e-6)


def test_tree_invariants():
    tree = seeded_tree(5, max_actions=3, max_transitions=2, transition_threshold=0.3, depth_bound=4,
                       depth_bound_lambda=lambda n: n.depth_bound - 1 - 2 * (random.random() < 0.5))
    tree.assert_index_is_tree()
    idx, val, ch = tree.index_tensor, tree.value_tensor, tree.chance_tensor
    # absorbing node and root conventions (reference tree.py:29-32, 338-349)
    assert int(idx[0].abs().sum()) == 0 and float(ch[0, 0, 0, 0]) == 1.0 and float(ch[0].sum()) == 1.0
    legal = tree.legal_tensor[:, 0]
    total = ch.sum(1)
    assert torch.allclose(total[legal != 0], torch.ones_like(total[legal != 0]), atol=1e-6)
    assert float(total[legal == 0].abs().sum()) == 0.0
    # expected value = sum_c chance * value ; value of a non-terminal child = its NE value
    close(tree.expected_value_tensor[:, 0], (ch * val).sum(1), atol=1e-6)
    nz = idx != 0
    close(val[nz], tree.root_value_tensor[idx[nz], 0], atol=0, rtol=0)


def test_nashconv_matches_reference_recursion(golden):
    from util.metric import NashConvData

    _, g = golden
    tree = tree_from_golden(g)
    net = mlp_from_golden(g, "learner")
    data = NashConvData(tree)
    data.get_nashconv_from_net(tree, net)
    close(data.joint_policy, g["nashconv.joint_policy"], atol=1e-6)
    close(data.row_best, g["nashconv.row_best"], atol=2e-6)
    close(data.col_best, g["nashconv.col_best"], atol=2e-6)
    close(data.reach_probability, g["nashconv.reach"], atol=1e-6)
    assert torch.equal(data.depth, t(g["nashconv.depth"]))
    close((data.row_best[1] + data.col_best[1]), g["nashconv.net"], atol=3e-6)


def test_nashconv_known_answers(golden):
    """The tree's own solution is unexploitable; the uniform policy is not (unless the game is trivial)."""
    from util.metric import NashConvData

    _, g = golden
    tree = tree_from_golden(g)
    data = NashConvData(tree)
    data.joint_policy = tree.solution_tensor.clone()
    data.get_nashconv(tree, tree.solution_tensor)
    assert abs((data.row_best[1] + data.col_best[1]).item()) <= 2e-6
    close((data.row_best[1] + data.col_best[1]), g["nashconv.solution"], atol=2e-6)
    a = tree.max_actions
    uniform = torch.cat([torch.nn.functional.normalize(tree.legal_tensor[:, 0, :, 0], p=1, dim=-1),
                         torch.nn.functional.normalize(tree.legal_tensor[:, 0, 0, :], p=1, dim=-1)], dim=1)
    data = NashConvData(tree)
    data.joint_policy = uniform.clone()
    data.get_nashconv(tree, uniform)
    assert (data.row_best[1] + data.col_best[1]).item() >= -1e-6
    assert uniform.shape == (tree.index_tensor.shape[0], 2 * a)


def test_reference_test_semantics():
    """What the reference's tests/test_nashconv.py asserts (exact zeros, reach sum 2) holds here too."""
    from util.metric import NashConvData

    for max_actions in range(2, 5):
        tree = seeded_tree(max_actions, max_actions=max_actions, max_transitions=1, depth_bound=3)
        data = NashConvData(tree)
        data.get_nashconv(tree, tree.solution_tensor)
        assert (data.row_best[1] + data.col_best[1]).item() == 0
        assert torch.sum(data.reach_probability).item() == 2


def test_tree_save_load_roundtrip(tmp_path, monkeypatch):
    from environment.tree import Tree

    tree = seeded_tree(2, max_actions=2, max_transitions=2, depth_bound=2)
    monkeypatch.setattr(Tree, "_saved_trees_dir", staticmethod(lambda: str(tmp_path)))
    tree.save("unit")
    other = Tree(max_actions=2, max_transitions=2)
    other.load("unit")
    assert other.hash == tree.hash
    assert torch.equal(other.index_tensor, tree.index_tensor)
    other.load()   # "recent"
    assert torch.equal(other.value_tensor, tree.value_tensor)


def test_library_exports_every_declared_symbol():
    import _b200

    header = open(os.path.join(REPO, "include", "rnad_b200.h")).read()
    declared = set(re.findall(r"RNAD_API\s+[\w\s\*]+?\b(rnad_\w+)\s*\(", header))
    assert declared == set(_b200.EXPORTS), declared ^ set(_b200.EXPORTS)
    assert os.path.exists(_b200.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_b200.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    L = _b200.lib()
    assert L.rnad_version() == 200
    evs, trs = ctypes.c_int(), ctypes.c_int()
    L.rnad_packed_strides(3, 2, ctypes.byref(evs), ctypes.byref(trs))
    assert (evs.value, trs.value) == (12, 8)
    with pytest.raises(_b200.RnadError):
        L.rnad_packed_strides(99, 2, ctypes.byref(evs), ctypes.byref(trs))
    assert L.rnad_rollout_tc_supported(3, 256) == 1 and L.rnad_rollout_tc_supported(6, 256) == 0


def test_product_path_fails_loudly_without_cuda(golden):
    """No CPU fallback: CPU tensors are refused by the kernel boundary."""
    import _b200
    from environment.episode import Episodes, States

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    _, g = golden
    tree = tree_from_golden(g)
    with pytest.raises(_b200.RnadError):
        States(tree, 4).observations()
    with pytest.raises(_b200.RnadError):
        Episodes(tree, 4).generate(mlp_from_golden(g, "net"))


def test_step_engine_fingerprint_sees_what_a_captured_step_bakes_in():
    """`LearnerStep.quick_key_of` (what RNaD compares every step instead of walking the modules) is stable while nothing
    changes and differs as soon as a hyper-parameter, the optimizer's settings, a net or a parameter tensor does."""
    import learn.fused as fused
    from learn.rnad import RNaD
    from nn.net import MLP

    tree = seeded_tree(3, max_actions=2, max_transitions=1, depth_bound=2)
    trial = RNaD(tree=tree, device=torch.device("cpu"), directory_name="pytest_quick_key", batch_size=64, eta=0.2, lr=1e-3,
                 net_params={"type": "MLP", "max_actions": 2, "width": 256})
    trial.net, trial.net_target, trial.net_reg, trial.net_reg_ = (MLP(2, 256) for _ in range(4))
    trial.optimizer = torch.optim.Adam(trial.net.parameters(), lr=1e-3, betas=(0.0, 0.999), eps=1e-8)
    key = fused.LearnerStep.quick_key_of
    k0 = key(trial)
    assert key(trial) == k0 and hash(k0) is not None
    trial.eta = 0.3
    k1 = key(trial)
    assert k1 != k0
    trial.optimizer.param_groups[0]["lr"] = 5e-4
    k2 = key(trial)
    assert k2 != k1
    trial.net_reg_ = MLP(2, 256)                                   # another module object
    k3 = key(trial)
    assert k3 != k2
    trial.net.value_fc0.weight.data = trial.net.value_fc0.weight.data.clone()     # same Parameter, another storage
    k4 = key(trial)
    assert k4 != k3
    trial.optimizer = torch.optim.Adam(trial.net.parameters(), lr=5e-4, betas=(0.0, 0.999), eps=1e-8)
    assert key(trial) != k4
