"""
Parity on the configurations bench.py measures (BASELINE.json configs 2-4), at their FULL batch sizes: the f16x2 (default)
rollout is launched exactly as benchmarked - cfg2: Tree(3, 2, depth_bound=4), 65,536 games = 256 tile pairs, most CTAs
playing two pairs; cfg3: max_transitions 3, depth 6, 14.9 M nodes (HBM-resident tables), 262,144 games; cfg4: A = 4,
T = 16, thinned ragged tree, 131,072 games (one GPU's share) - and a random sample of its games is replayed on the
CPU oracle: node ids, masks, observations, rewards bit-exact, every sampled action / chance outcome the inverse-CDF
choice at the Philox uniform of (seed, game, half-move), policy / value within the engine's tolerance.  Whole-batch
properties that need no replay are checked on every game: terminal at the end, policy rows sum to one, illegal
actions have probability zero, rewards only on final column half-moves, t_eff.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_env_rollout import TOL, check_rollout_against_oracle  # noqa: E402

pytestmark = pytest.mark.gpu
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, REPO)


def _bench():
    import bench

    return bench


def _tables(tree):
    return {"index": tree.index_tensor.cpu(), "value": tree.value_tensor.cpu(), "chance": tree.chance_tensor.cpu(),
            "expected_value": tree.expected_value_tensor.cpu(), "legal": tree.legal_tensor.cpu()}


def _whole_batch_properties(ep, t_max, regular):
    idx, pol, masks, rew, turns = ep.indices, ep.policy, ep.masks, ep.rewards, ep.turns
    valid = idx != 0
    assert ep.t_eff + 1 <= t_max
    if regular:
        assert ep.t_eff + 1 == t_max and bool(valid.all()), "a regular tree keeps every game alive for 2 * depth half-moves"
    assert bool((idx[0] == 1).all())
    assert bool(((pol.sum(-1) - 1).abs() < 1e-5).all())
    assert bool((pol[masks == 0] == 0).all())
    assert bool((turns == (torch.arange(idx.shape[0], device=idx.device) & 1)[:, None]).all())
    assert bool((rew[0::2] == 0).all()), "rewards arrive on column half-moves only"
    last = valid.sum(0) - 1                                  # a game's last live half-move is a column half-move
    assert bool((last % 2 == 1).all())
    got = rew.gather(0, last[None])[0]
    assert bool((rew.abs().sum(0) == got.abs()).all()), "a game is paid once, when it ends"
    assert bool((ep.actions.sum(-1) == 1).all()) and bool((ep.actions * (1 - masks) == 0).all())


@pytest.mark.parametrize("config,sample", [("cfg2", 16384), ("cfg3", 12000), ("cfg4", 12000)])
def test_benchmarked_rollout_replays_on_the_oracle(config, sample):
    from environment.episode import Episodes
    from nn.net import MLP

    bench = _bench()
    depth, a, c, batch = bench.CONFIGS[config]
    dev = torch.device("cuda")
    if config in bench.FAST_TREE_CONFIGS:
        tree = bench.fast_tree(config, depth, a, c, dev)
    else:
        tree = bench.make_tree(depth, a, c, seed=0)
        tree.to(dev)
    torch.manual_seed(1234)
    net = MLP(a, 256, device=dev)                 # the benchmark's net
    w = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    torch.manual_seed(99)
    ep = Episodes(tree, batch)
    ep.generate(net)
    assert ep.precision == "f16x2"
    t_max = tree.packed().max_half_moves
    assert t_max == 2 * depth
    _whole_batch_properties(ep, t_max, regular=config != "cfg4")
    if config == "cfg4":
        frac = float((ep.indices != 0).float().mean()) * (ep.t_eff + 1) / t_max
        assert 0.3 < frac < 0.9, frac                # ragged: about half of the (t, b) slots are valid
    rng = np.random.default_rng(5)
    # whole tiles from both ends and the middle (first / last CTA, a CTA's second tile pair) plus a random scatter
    tiles = np.concatenate([np.arange(0, 256), np.arange(batch - 384, batch), np.arange(batch // 2 + 64, batch // 2 + 448)])
    games = np.unique(np.concatenate([tiles, rng.choice(batch, size=sample, replace=False)]))
    tables = _tables(tree)
    check_rollout_against_oracle(ep, tables, w, seed=ep.states.seed, tol=TOL["f16x2"], precision="f16x2", games=games)
