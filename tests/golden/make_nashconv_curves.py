"""
Reference NashConv-vs-steps curves (BASELINE.json config 5: random-tree sweep, 10 seeds x depth_bound 3..8): runs the
UNMODIFIED reference (/root/reference) `RNaD.run` on CPU on seeded random trees with main.py's tree and learner
settings (main.py:31-81) and records the NashConv of the target net along the run.  Build container only; writes
tests/golden/nashconv_curves.json.

    python tests/golden/make_nashconv_curves.py [n_seeds] [n_updates] [workers] [depths, e.g. 3,4,5,6,7,8]

One (depth, seed) job per subprocess, `workers` of them at a time, one torch thread each (the jobs are Python-bound:
tree generation is ~10 ms/node, one NashConv evaluation ~0.8 ms/node).  Deep trees are evaluated at fewer points
(every EVAL_EVERY[depth] updates of delta_m = 100 learner steps).

Harness-side changes only (as in make_golden.py): stand-in pygambit on sys.path, the reference imported from a
scratch copy under /tmp (it writes saved_runs/ next to itself), b1_adam=0.0 (torch 2.11), seeds set from outside.
"""
import json
import os
import random
import shutil
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
DELTA_M = 100
EVAL_EVERY = {3: 1, 4: 1, 5: 1, 6: 2, 7: 5, 8: 5}
PARTS = os.path.join(HERE, "curve_parts")


def config(n_updates):
    return {"max_actions": 3, "max_transitions": 2, "transition_threshold": 0.3, "eta": 0.2, "lr": 1e-3,
            "gamma_averaging": 0.01, "batch_size": 512, "logit_clip": 2, "delta_m": DELTA_M, "updates": n_updates,
            "width": 256, "depth_lambda": "depth_bound - 1 - 2 * (random() < 0.5)",
            "eval_every": {str(k): v for k, v in EVAL_EVERY.items()}}


def job(depth, seed_idx, n_updates):
    import numpy as np
    import torch

    sys.path.insert(0, os.path.join(HERE, "_standin"))
    scratch = tempfile.mkdtemp(prefix="rnad_ref_curves_")
    ref = os.path.join(scratch, "ref")
    shutil.copytree("/root/reference", ref)
    sys.path.insert(0, ref)
    from environment.tree import Tree
    from learn.rnad import RNaD
    from util.metric import NashConvData

    torch.set_num_threads(1)
    sys.setrecursionlimit(100000)

    def nashconv(tree, net):
        data = NashConvData(tree)
        data.get_nashconv_from_net(tree, net)
        return float(data.row_best[1] + data.col_best[1])

    seed = 1000 * depth + seed_idx
    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)
    t0 = time.time()
    tree = Tree(device=torch.device("cpu"), max_actions=3, max_transitions=2, transition_threshold=0.3,
                depth_bound=depth, depth_bound_lambda=lambda t: t.depth_bound - 1 - 2 * (random.random() < 0.5))
    tree.generate()
    t_gen = time.time() - t0
    trial = RNaD(tree=tree, device=torch.device("cpu"), directory_name=f"curve_d{depth}_s{seed_idx}_{os.getpid()}",
                 eta=0.2, bounds=[n_updates], delta_m=[DELTA_M], lr=1e-3, gamma_averaging=0.01, batch_size=512,
                 logit_clip=2, b1_adam=0.0, net_params={"type": "MLP", "max_actions": 3, "width": 256}, wandb=False)
    trial._RNaD__initialize()
    every = EVAL_EVERY[depth]
    updates, curve = [0], [nashconv(tree, trial.net_target)]
    for m in range(n_updates):
        trial.bounds = [m + 1]
        trial._RNaD__resume(checkpoint_mod=10 ** 9, expl_mod=10 ** 9, log_mod=10 ** 9)
        if (m + 1) % every == 0 or m + 1 == n_updates:
            updates.append(m + 1)
            curve.append(nashconv(tree, trial.net_target))
    rec = {"depth": depth, "seed": seed, "nodes": int(tree.index_tensor.shape[0]), "updates": updates,
           "nashconv": curve, "generate_s": round(t_gen, 1), "total_s": round(time.time() - t0, 1)}
    os.makedirs(PARTS, exist_ok=True)
    with open(os.path.join(PARTS, f"d{depth}_s{seed_idx}.json"), "w") as f:
        json.dump(rec, f)
    shutil.rmtree(scratch, ignore_errors=True)
    print(f"depth {depth} seed {seed_idx}: {rec['nodes']} nodes, NashConv {curve[0]:.3f} -> {curve[-1]:.3f} "
          f"(generate {t_gen:.0f} s, total {rec['total_s']:.0f} s)", flush=True)


def merge(n_updates):
    out = {"config": config(n_updates), "curves": {}}
    for name in sorted(os.listdir(PARTS), key=lambda s: (int(s[1:s.index("_")]), int(s[s.index("_s") + 2:-5]))):
        out["curves"][name[:-5]] = json.load(open(os.path.join(PARTS, name)))
    with open(os.path.join(HERE, "nashconv_curves.json"), "w") as f:
        json.dump(out, f, indent=1)
    return out


if __name__ == "__main__":
    if sys.argv[1] == "--job":
        job(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
        sys.exit(0)
    if sys.argv[1] == "--merge":
        print(len(merge(int(sys.argv[2]))["curves"]), "curves merged")
        sys.exit(0)
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    n_updates = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    workers = int(sys.argv[3]) if len(sys.argv) > 3 else max(1, (os.cpu_count() or 2) - 2)
    depths = [int(d) for d in sys.argv[4].split(",")] if len(sys.argv) > 4 else [3, 4, 5, 6, 7, 8]
    jobs = [(d, s) for d in sorted(depths, reverse=True) for s in range(n_seeds)      # longest first
            if not os.path.exists(os.path.join(PARTS, f"d{d}_s{s}.json"))]

    def run(ds):
        subprocess.run([sys.executable, os.path.abspath(__file__), "--job", str(ds[0]), str(ds[1]), str(n_updates)],
                       check=False)

    with ThreadPoolExecutor(workers) as pool:
        list(pool.map(run, jobs))
    print(len(merge(n_updates)["curves"]), "curves merged")
