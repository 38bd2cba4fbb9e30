"""
Level-synchronous tree generator (environment/fast_tree.py, SURVEY.md 8 f#1): structural invariants of the
reference's tables (tree.py:14-63, 368-383), the batched solver against the host solver, NashConv of the stored
solution, and - on the GPU - the fused rollout on a generated tree replayed by the CPU oracle.
"""
import numpy as np
import pytest
import torch

from environment.fast_tree import depth_jitter, solve_zero_sum_batched
from environment.tree import Tree
from util.matrix_game import solve_zero_sum
from util.metric import NashConvData


def check_tables(tree):
    tree.assert_index_is_tree()
    a, c = tree.max_actions, tree.max_transitions
    S = tree.index_tensor.shape[0]
    assert tree.index_tensor.shape == (S, c, a, a) and tree.index_tensor.dtype == torch.int64
    assert tree.value_tensor.shape == (S, c, a, a) and tree.chance_tensor.shape == (S, c, a, a)
    assert tree.expected_value_tensor.shape == (S, 1, a, a) and tree.legal_tensor.shape == (S, 1, a, a)
    assert tree.root_value_tensor.shape == (S, 1) and tree.solution_tensor.shape == (S, 2 * a)
    legal = tree.legal_tensor[:, 0]
    rows, cols = legal[:, :, 0].sum(1).long(), legal[:, 0, :].sum(1).long()
    ar = torch.arange(a, device=legal.device)
    rect = ((ar.view(1, a, 1) < rows.view(S, 1, 1)) & (ar.view(1, 1, a) < cols.view(S, 1, 1))).float()
    assert torch.equal(legal, rect), "legal masks are prefix rectangles (tree.py:133)"
    total = tree.chance_tensor.sum(1)
    assert torch.allclose(total[legal > 0], torch.ones(1, device=legal.device), atol=1e-6)
    assert float(total[legal == 0].abs().sum()) == 0
    # absorbing node (tree.py:338-349)
    assert int(tree.index_tensor[0].abs().sum()) == 0 and float(tree.chance_tensor[0].sum()) == 1.0
    assert float(tree.chance_tensor[0, 0, 0, 0]) == 1.0 and float(tree.legal_tensor[0].sum()) == 1.0
    ev = (tree.value_tensor * tree.chance_tensor).sum(1)
    assert torch.equal(ev, tree.expected_value_tensor[:, 0])                      # tree.py:280-282
    idx = tree.index_tensor
    assert torch.equal(tree.value_tensor[idx > 0], tree.root_value_tensor[idx[idx > 0], 0])   # tree.py:269-272
    leaf = (idx == 0) & (tree.chance_tensor > 0)
    leaf[0] = False
    assert bool(torch.isin(tree.value_tensor[leaf], torch.tensor(tree.terminal_values, dtype=torch.float32,
                                                                 device=legal.device)).all())
    assert bool((tree.index_tensor[tree.chance_tensor == 0] == 0).all())
    sol = tree.solution_tensor[1:]
    assert torch.allclose(sol[:, :a].sum(1), torch.ones(1, device=legal.device), atol=1e-6)
    assert torch.allclose(sol[:, a:].sum(1), torch.ones(1, device=legal.device), atol=1e-6)


def test_batched_solver_matches_host_solver():
    gen = torch.Generator().manual_seed(0)
    n, a = 1500, 4
    M = torch.rand(n, a, a, generator=gen, dtype=torch.float64) * 2 - 1
    M[::5] = torch.round(M[::5] * 2) / 2                  # ties and degenerate games
    M[1::11] = M[1::11, :1, :1]                           # constant matrices
    rows = torch.randint(1, a + 1, (n,), generator=gen)
    cols = torch.randint(1, a + 1, (n,), generator=gen)
    x, y, v, failed = solve_zero_sum_batched(M, rows, cols)
    assert not bool(failed.any())
    for i in range(n):
        r, c = int(rows[i]), int(cols[i])
        xs, ys, vs = solve_zero_sum(M[i, :r, :c].numpy())
        assert abs(vs - float(v[i])) < 1e-12
        np.testing.assert_allclose(x[i, :r].numpy(), xs, atol=1e-12)
        np.testing.assert_allclose(y[i, :c].numpy(), ys, atol=1e-12)
        assert float(x[i, r:].abs().sum()) == 0 and float(y[i, c:].abs().sum()) == 0


@pytest.mark.parametrize("a,c,depth,thr,jitter", [(2, 1, 2, 0.0, None), (3, 2, 4, 0.0, None), (3, 3, 3, 0.2, None),
                                                  (4, 2, 5, 0.3, 0.5), (3, 2, 6, 0.3, 0.5)])
def test_generate_fast_tables_and_equilibrium(a, c, depth, thr, jitter):
    tree = Tree(max_actions=a, max_transitions=c, depth_bound=depth, transition_threshold=thr)
    tree.generate_fast(seed=depth, child_spec=depth_jitter(jitter) if jitter else None)
    check_tables(tree)
    if jitter is None and thr == 0.0:
        k = a * a * c                                      # full regular tree
        assert tree.index_tensor.shape[0] == 1 + sum(k ** j for j in range(depth))
    data = NashConvData(tree)
    data.joint_policy = tree.solution_tensor.clone()
    data.get_nashconv(tree, tree.solution_tensor)
    assert float(data.row_best[1] + data.col_best[1]) < 1e-6     # the stored solution is an equilibrium of the whole tree
    uniform = torch.nn.functional.normalize(torch.cat([tree.legal_tensor[:, 0, :, 0], tree.legal_tensor[:, 0, 0, :]], 1)
                                            .view(-1, 2, a), p=1, dim=-1).view(-1, 2 * a)
    data = NashConvData(tree)
    data.joint_policy = uniform.clone()
    data.get_nashconv(tree, uniform)
    if depth >= 3:
        assert float(data.row_best[1] + data.col_best[1]) > 1e-3
    # same seed, same tree
    again = Tree(max_actions=a, max_transitions=c, depth_bound=depth, transition_threshold=thr)
    again.generate_fast(seed=depth, child_spec=depth_jitter(jitter) if jitter else None)
    assert torch.equal(again.index_tensor, tree.index_tensor) and torch.equal(again.value_tensor, tree.value_tensor)


@pytest.mark.gpu
def test_generate_fast_on_gpu_and_rollout_replay():
    from environment.episode import Episodes
    from oracle import rnad_oracle as orc
    from test_gpu_env_rollout import check_rollout_against_oracle, wide_net, TOL

    dev = torch.device("cuda")
    tree = Tree(device=dev, max_actions=3, max_transitions=2, depth_bound=6, transition_threshold=0.3)
    tree.generate_fast(seed=3, child_spec=depth_jitter(0.4))
    check_tables(tree)
    assert tree.index_tensor.is_cuda and tree.index_tensor.shape[0] > 20000
    tables = {"index": tree.index_tensor.cpu(), "value": tree.value_tensor.cpu(), "chance": tree.chance_tensor.cpu(),
              "expected_value": tree.expected_value_tensor.cpu(), "legal": tree.legal_tensor.cpu()}
    net, w = wide_net(3, 5, "cuda")
    torch.manual_seed(11)
    ep = Episodes(tree, 30000)
    ep.generate(net)
    assert ep.t_eff + 1 <= 12
    check_rollout_against_oracle(ep, tables, w, seed=ep.states.seed, tol=TOL[ep.precision], precision=ep.precision)
