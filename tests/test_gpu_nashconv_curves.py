"""
NashConv-vs-steps curves against the reference (BASELINE.json config 5, reduced).

tests/golden/nashconv_curves.json holds the curves of the UNMODIFIED reference (`RNaD.run` on CPU, main.py's tree
and learner settings; tests/golden/make_nashconv_curves.py) on ten seeded random trees.  Here the same trees are
rebuilt from the same seeds (Tree.generate follows the reference's RNG draws), trained with this repository's RNaD
on the GPU with the same hyper-parameters, and the curves are compared statistically: the sampling streams differ
(Philox inverse-CDF vs torch.multinomial), so trajectories - and single curves - cannot match, their means do.
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

    ours = {}
    for name, rec in ref["curves"].items():
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
        ours[name] = curve

    for depth in sorted({r["depth"] for r in ref["curves"].values()}):
        names = [n for n, r in ref["curves"].items() if r["depth"] == depth]
        r = np.array([ref["curves"][n]["nashconv"] for n in names])
        o = np.array([ours[n] for n in names])
        rm, om, rs = r.mean(0), o.mean(0), r.std(0)
        print(f"depth {depth}: reference mean {np.round(rm[[0, 1, 2, 5, 10, 20]], 3)}  ours {np.round(om[[0, 1, 2, 5, 10, 20]], 3)}")
        # same start (same trees; initial nets are untrained), same decay, same plateau
        assert abs(om[0] - rm[0]) < 0.25
        tail_r, tail_o = r[:, -5:].mean(), o[:, -5:].mean()
        assert abs(tail_o - tail_r) < max(0.15, 2.5 * r[:, -5:].mean(1).std() / np.sqrt(len(names)) + 0.1), (tail_o, tail_r)
        assert np.abs(om - rm).mean() < 0.2, np.abs(om - rm).mean()
        assert om[-1] < 0.75 * om[0]
        assert tail_o < 0.8 * om[0] and tail_r < 0.8 * rm[0]
