#!/usr/bin/env python
"""
BASELINE.json config 5 on the GPU: the random-tree sweep (10 seeds x depth_bound 3..8, reference main.py's tree and
learner settings, main.py:31-81) trained with THIS repository's RNaD, NashConv of the target net evaluated at the
update counts the reference curves were recorded at (tests/golden/nashconv_curves.json, written by
tests/golden/make_nashconv_curves.py from the unmodified reference).  Every tree is rebuilt from its seed with
Tree.generate (which follows the reference's RNG draws; the node count is checked against the reference's).

    python scripts/cfg5_sweep.py [--engines default,fp32] [--depths 3,4,5,6,7,8] [--seeds 10] [--out gpurun_out/cfg5.json]

engines: "default" = fused tensor-core kernels (tf32 rollout, tf32 learner passes, LearnerStep graph);
         "fp32"    = fp32 rollout engine + reference-style fp32 torch GEMM / autograd learner (`learner_engine="torch"`).
Prints mean +- sd per depth at every recorded update for the reference and for each engine.
"""
import argparse
import json
import os
import random
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "r-nad_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def nashconv(tree, net):
    from util.metric import NashConvData

    data = NashConvData(tree)
    data.get_nashconv_from_net(tree, net)
    return float(data.row_best[1] + data.col_best[1])


TREE_KEYS = ("index_tensor", "value_tensor", "chance_tensor", "expected_value_tensor", "legal_tensor",
             "root_value_tensor", "solution_tensor")


def build_tree(job):
    """(worker process, CPU only) the reference's tree for (depth, seed), rebuilt from the seed; saved to a file."""
    name, rec, cfg, path = job
    from environment.tree import Tree

    torch.set_num_threads(1)
    seed = rec["seed"]
    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)
    sys.setrecursionlimit(100000)
    tree = Tree(device=torch.device("cpu"), max_actions=cfg["max_actions"], max_transitions=cfg["max_transitions"],
                transition_threshold=cfg["transition_threshold"], depth_bound=rec["depth"],
                depth_bound_lambda=lambda t: t.depth_bound - 1 - 2 * (random.random() < 0.5))
    tree.generate()
    assert int(tree.index_tensor.shape[0]) == rec["nodes"], "the seeded tree differs from the reference's"
    torch.save({k: getattr(tree, k) for k in TREE_KEYS} | {"hash": tree.hash}, path)
    return name


def run_curve(rec, cfg, engine, dev, tree_file):
    from environment.tree import Tree
    from learn.rnad import RNaD

    seed = rec["seed"]
    torch.manual_seed(seed)
    tree = Tree(device=torch.device("cpu"), max_actions=cfg["max_actions"], max_transitions=cfg["max_transitions"],
                transition_threshold=cfg["transition_threshold"], depth_bound=rec["depth"])
    for key, value in torch.load(tree_file).items():
        setattr(tree, key, value)
    tree.to(dev)
    trial = RNaD(tree=tree, device=dev, directory_name=f"cfg5_{engine}_d{rec['depth']}_s{seed}_{os.getpid()}", eta=cfg["eta"],
                 bounds=[cfg["updates"]], delta_m=[cfg["delta_m"]], lr=cfg["lr"], gamma_averaging=cfg["gamma_averaging"],
                 batch_size=cfg["batch_size"], logit_clip=cfg["logit_clip"],
                 net_params={"type": "MLP", "max_actions": cfg["max_actions"], "width": cfg["width"]})
    if engine == "fp32":
        trial.learner_engine = "torch"
        os.environ["RNAD_ROLLOUT_PRECISION"] = "fp32"
    else:
        os.environ.pop("RNAD_ROLLOUT_PRECISION", None)
    trial._RNaD__initialize()
    wanted = set(rec["updates"])
    curve = {0: nashconv(tree, trial.net_target)}
    for m in range(cfg["updates"]):
        trial.bounds = [m + 1]
        trial._RNaD__resume(checkpoint_mod=10 ** 9, expl_mod=10 ** 9, log_mod=10 ** 9)
        if m + 1 in wanted:
            curve[m + 1] = nashconv(tree, trial.net_target)
    import shutil

    shutil.rmtree(trial.directory, ignore_errors=True)
    return [curve[u] for u in rec["updates"]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engines", default="default,fp32")
    ap.add_argument("--depths", default="3,4,5,6,7,8")
    ap.add_argument("--seeds", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "cfg5_sweep.json"))
    args = ap.parse_args()
    ref = json.load(open(os.path.join(REPO, "tests", "golden", "nashconv_curves.json")))
    cfg = ref["config"]
    depths = [int(d) for d in args.depths.split(",")]
    dev = torch.device("cuda")
    out = {"config": cfg, "engines": {}, "reference": {}}
    # phase 1, host cores only: the 60 trees, rebuilt from their seeds in parallel (0.3 ms per node and process)
    import multiprocessing as mp
    import tempfile

    tmp = tempfile.mkdtemp(prefix="cfg5_trees_")
    jobs = [(name, rec, cfg, os.path.join(tmp, name + ".pt")) for name, rec in ref["curves"].items()
            if rec["depth"] in depths and int(name.split("_s")[1]) < args.seeds]
    jobs.sort(key=lambda j: -j[1]["nodes"])
    t0 = time.time()
    with mp.get_context("spawn").Pool(min(len(jobs), max(1, (os.cpu_count() or 2) - 1))) as pool:
        pool.map(build_tree, jobs, chunksize=1)
    tree_files = {name: path for name, _, _, path in jobs}
    print(f"{len(jobs)} trees rebuilt in {time.time() - t0:.0f} s", flush=True)
    for engine in args.engines.split(","):
        out["engines"][engine] = {}
        for name, rec in ref["curves"].items():
            if name not in tree_files:
                continue
            t0 = time.time()
            curve = run_curve(rec, cfg, engine, dev, tree_files[name])
            out["engines"][engine][name] = curve
            out["reference"][name] = {"updates": rec["updates"], "nashconv": rec["nashconv"], "nodes": rec["nodes"], "depth": rec["depth"]}
            print(f"{engine} {name}: {rec['nodes']} nodes, NashConv {curve[0]:.3f} -> {curve[-1]:.3f} "
                  f"(reference {rec['nashconv'][0]:.3f} -> {rec['nashconv'][-1]:.3f}), {time.time() - t0:.0f} s", flush=True)
            with open(args.out, "w") as f:
                json.dump(out, f)
    # summary
    print("\\ndepth | updates | reference mean+-sd | " + " | ".join(args.engines.split(",")))
    for depth in depths:
        names = [n for n, r in out["reference"].items() if r["depth"] == depth]
        if not names:
            continue
        ups = out["reference"][names[0]]["updates"]
        r = np.array([out["reference"][n]["nashconv"] for n in names])
        for j in sorted(set([0, len(ups) // 4, len(ups) // 2, len(ups) - 1])):
            cells = [f"{r[:, j].mean():.3f}+-{r[:, j].std():.3f}"]
            for engine in out["engines"]:
                o = np.array([out["engines"][engine][n] for n in names])
                cells.append(f"{o[:, j].mean():.3f}+-{o[:, j].std():.3f}")
            print(f"{depth} | {ups[j] * cfg['delta_m']:5d} steps | " + " | ".join(cells))


if __name__ == "__main__":
    main()
