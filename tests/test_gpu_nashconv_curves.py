"""
NashConv-vs-steps curves against the reference (BASELINE.json config 5, reduced).

tests/golden/nashconv_curves.json holds the curves of the UNMODIFIED reference (`RNaD.run` on CPU, main.py's tree
and learner settings; tests/golden/make_nashconv_curves.py) on seeded random trees: 10 seeds x depth_bound 3..8.
Here the trees of depth_bound 3 and 4 (20 trees; the whole sweep is scripts/cfg5_sweep.py, its results are kept under
profiles/) are rebuilt from the same seeds (Tree.generate follows the reference's RNG draws), trained with this
repository's RNaD on the GPU with the same hyper-parameters, and the curves are compared statistically: the sampling
streams differ (Philox inverse-CDF vs torch.multinomial), so trajectories - and single curves - cannot match, their
means do.  Bounds: the difference of the seed-averaged plateaus must be within 3 standard errors of the difference
(both sets of curves have seed-to-seed spread) plus 0.03.
"""
import json
import os
import random

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CURVES = os.path.join(HERE, "golden", "nashconv_curves.json")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(CURVES), reason="reference curves not generated")
def test_nashconv_curves_track_the_reference():
    from environment.tree import Tree
    from learn.rnad import RNaD
    from util.metric import NashConvData

    ref = json.load(open(CURVES))
    cfg = ref["config"]
    dev = torch.device("cuda")

    def nashconv(tree, net):
        data = NashConvData(tree)
        data.get_nashconv_from_net(tree, net)
        return float(data.row_best[1] + data.col_best[1])

    depths = (3, 4)
    ours = {}
    for name, rec in ref["curves"].items():
        if rec["depth"] not in depths:
            continue
        seed = rec["seed"]
        np.random.seed(seed)
        random.seed(seed)
        torch.manual_seed(seed)
        tree = Tree(device=torch.device("cpu"), max_actions=cfg["max_actions"], max_transitions=cfg["max_transitions"],
                    transition_threshold=cfg["transition_threshold"], depth_bound=rec["depth"],
                    depth_bound_lambda=lambda t: t.depth_bound - 1 - 2 * (random.random() < 0.5))
        tree.generate()
        assert int(tree.index_tensor.shape[0]) == rec["nodes"], "the seeded tree differs from the reference's"
        tree.to(dev)
        trial = RNaD(tree=tree, device=dev, directory_name=f"pytest_curve_{name}_{os.getpid()}", eta=cfg["eta"],
                     bounds=[cfg["updates"]], delta_m=[cfg["delta_m"]], lr=cfg["lr"], gamma_averaging=cfg["gamma_averaging"],
                     batch_size=cfg["batch_size"], logit_clip=cfg["logit_clip"],
                     net_params={"type": "MLP", "max_actions": cfg["max_actions"], "width": cfg["width"]})
        trial._RNaD__initialize()
        curve = [nashconv(tree, trial.net_target)]
        trial.run(checkpoint_mod=10 ** 9, expl_mod=1, log_mod=10 ** 9)
        curve += [v for _, v in trial.nashconv_history] + [nashconv(tree, trial.net_target)]
        assert len(curve) == cfg["updates"] + 1
        ours[name] = [curve[u] for u in rec.get("updates", range(len(curve)))]   # the update counts the reference recorded

    for depth in depths:
        names = [n for n, r in ref["curves"].items() if r["depth"] == depth]
        r = np.array([ref["curves"][n]["nashconv"] for n in names])
        o = np.array([ours[n] for n in names])
        rm, om = r.mean(0), o.mean(0)
        show = [0, 1, 2, len(rm) // 4, len(rm) // 2, len(rm) - 1]
        print(f"depth {depth}: reference mean {np.round(rm[show], 3)}  ours {np.round(om[show], 3)}")
        # same start (same trees; initial nets are untrained), same decay, same plateau
        assert abs(om[0] - rm[0]) < 0.25
        tail_r, tail_o = r[:, -5:].mean(1), o[:, -5:].mean(1)
        stderr = np.sqrt(tail_r.var(ddof=1) / len(names) + tail_o.var(ddof=1) / len(names))
        assert abs(tail_o.mean() - tail_r.mean()) < 3 * stderr + 0.03, (tail_o.mean(), tail_r.mean(), stderr)
        assert np.abs(om - rm).mean() < 0.15, np.abs(om - rm).mean()
        assert om[-1] < 0.75 * om[0]
        assert tail_o.mean() < 0.8 * om[0] and tail_r.mean() < 0.8 * rm[0]
