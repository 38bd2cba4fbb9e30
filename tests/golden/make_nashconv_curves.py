"""
Reference NashConv-vs-steps curves (BASELINE.json config 5, reduced): runs the UNMODIFIED reference
(/root/reference) `RNaD.run` on CPU on seeded random trees with main.py's tree and learner settings and records
the NashConv of the target net after every update.  Build container only; writes tests/golden/nashconv_curves.json.

    python tests/golden/make_nashconv_curves.py [n_seeds] [n_updates]

Harness-side changes only (as in make_golden.py): stand-in pygambit on sys.path, the reference imported from a
scratch copy under /tmp (it writes saved_runs/ next to itself), b1_adam=0.0 (torch 2.11), seeds set from outside.
"""
import json
import os
import random
import shutil
import sys
import tempfile
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_standin"))
SCRATCH = tempfile.mkdtemp(prefix="rnad_ref_curves_")
REF = os.path.join(SCRATCH, "ref")
shutil.copytree("/root/reference", REF)
sys.path.insert(0, REF)

from environment.tree import Tree  # noqa: E402
from learn.rnad import RNaD  # noqa: E402
from util.metric import NashConvData  # noqa: E402

N_SEEDS = int(sys.argv[1]) if len(sys.argv) > 1 else 5
N_UPDATES = int(sys.argv[2]) if len(sys.argv) > 2 else 20
DELTA_M = 100
CONFIG = {"max_actions": 3, "max_transitions": 2, "transition_threshold": 0.3, "eta": 0.2, "lr": 1e-3,
          "gamma_averaging": 0.01, "batch_size": 512, "logit_clip": 2, "delta_m": DELTA_M, "updates": N_UPDATES,
          "width": 256, "depth_lambda": "depth_bound - 1 - 2 * (random() < 0.5)"}


def seed_all(s):
    np.random.seed(s)
    random.seed(s)
    torch.manual_seed(s)


def nashconv(tree, net):
    data = NashConvData(tree)
    data.get_nashconv_from_net(tree, net)
    return float(data.row_best[1] + data.col_best[1])


out = {"config": CONFIG, "curves": {}}
torch.set_num_threads(os.cpu_count())
for depth in (3, 4):
    for seed in range(N_SEEDS):
        seed_all(1000 * depth + seed)
        tree = Tree(device=torch.device("cpu"), max_actions=3, max_transitions=2, transition_threshold=0.3,
                    depth_bound=depth,
                    depth_bound_lambda=lambda t: t.depth_bound - 1 - 2 * (random.random() < 0.5))
        tree.generate()
        t0 = time.time()
        trial = RNaD(tree=tree, device=torch.device("cpu"), directory_name=f"curve_d{depth}_s{seed}_{os.getpid()}",
                     eta=0.2, bounds=[N_UPDATES], delta_m=[DELTA_M], lr=1e-3, gamma_averaging=0.01, batch_size=512,
                     logit_clip=2, b1_adam=0.0, net_params={"type": "MLP", "max_actions": 3, "width": 256}, wandb=False)
        trial._RNaD__initialize()
        curve = [nashconv(tree, trial.net_target)]
        for m in range(N_UPDATES):
            trial.bounds = [m + 1]
            trial._RNaD__resume(checkpoint_mod=10 ** 9, expl_mod=10 ** 9, log_mod=10 ** 9)
            curve.append(nashconv(tree, trial.net_target))
        out["curves"][f"d{depth}_s{seed}"] = {"depth": depth, "seed": 1000 * depth + seed,
                                              "nodes": int(tree.index_tensor.shape[0]), "nashconv": curve}
        print(f"depth {depth} seed {seed}: {tree.index_tensor.shape[0]} nodes, NashConv {curve[0]:.3f} -> {curve[-1]:.3f} "
              f"({time.time() - t0:.0f} s)", flush=True)
        with open(os.path.join(HERE, "nashconv_curves.json"), "w") as f:
            json.dump(out, f, indent=1)
shutil.rmtree(SCRATCH, ignore_errors=True)
