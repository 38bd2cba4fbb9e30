"""
Driver: generate (or load) a stochastic matrix tree, then run R-NaD on it for a
sweep of regularisation strengths and log the exploitability of the target net.

Same flow as the reference's main.py (:29-81) - which also runs unchanged against
these packages, see INTEGRATION.md - with the BASELINE.json configurations as
presets and one process per GPU when launched under torchrun:

    python main.py                               # reference defaults (3x3, depth <= 4, B=512)
    python main.py --preset cfg2 --etas 0.2      # depth 4, branching 3, chance 2, B=65536
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 main.py --preset cfg2

Under torchrun every rank must hold the SAME tree and write into the SAME run directories: the tree seed (drawn by
rank 0 when --seed is not given) and the timestamp of the directory names are broadcast from rank 0, and
`RNaD.__initialize` asserts that the trees' hashes agree.
"""

import argparse
import logging
import os
from random import random
from time import time

import torch

from environment.tree import Tree
from learn.rnad import RNaD

PRESETS = {
    # reference main.py:31-39
    "main": dict(max_actions=3, max_transitions=2, transition_threshold=0.3, depth_bound=4, ragged=True, batch=2 ** 9),
    "cfg1": dict(max_actions=2, max_transitions=1, transition_threshold=0.0, depth_bound=2, ragged=False, batch=256),
    "cfg2": dict(max_actions=3, max_transitions=2, transition_threshold=0.0, depth_bound=4, ragged=False, batch=65536),
}


def build_tree(preset, device, load=None):
    p = PRESETS[preset]
    kwargs = dict(device=torch.device("cpu"), max_actions=p["max_actions"], max_transitions=p["max_transitions"],
                  transition_threshold=p["transition_threshold"], depth_bound=p["depth_bound"],
                  desc=f"{preset}: {p['max_actions']}x{p['max_actions']} stochastic tree, depth up to {p['depth_bound']}")
    if p["ragged"]:
        kwargs["depth_bound_lambda"] = lambda node: node.depth_bound - 1 - 2 * (random() < 0.5)
    tree = Tree(**kwargs)
    if load:
        tree.load(load)
    else:
        tree.generate()
        tree.assert_index_is_tree()
    tree.to(device)
    return tree


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="main", choices=sorted(PRESETS))
    ap.add_argument("--etas", type=float, nargs="+", default=[0, 0.2, 0.5, 1])
    ap.add_argument("--updates", type=int, default=64, help="number of regularisation updates (bounds)")
    ap.add_argument("--delta-m", type=int, default=100, help="learner steps per update")
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--load-tree", default=None, help="saved_trees/<name> to load instead of generating")
    ap.add_argument("--save-tree", default="small_tree")
    ap.add_argument("--seed", type=int, default=None)
    args = ap.parse_args()

    logging.basicConfig(level=logging.DEBUG)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("the R-NaD hot path runs on CUDA kernels only (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    import learn.dp as dp

    seed = args.seed
    if world > 1 and seed is None:
        seed = int.from_bytes(os.urandom(4), "little")    # rank 0's draw wins: the ranks must build the same tree
    seed, timestamp = dp.broadcast_object((seed, str(int(time()))))
    if seed is not None:
        import numpy as np
        import random as pyrandom

        np.random.seed(seed)
        pyrandom.seed(seed)
        torch.manual_seed(seed)               # same tree on every rank

    tree = build_tree(args.preset, device, load=args.load_tree)
    if seed is not None:
        torch.manual_seed(seed * 1000 + int(os.environ.get("RANK", "0")))   # different games per rank
    if int(os.environ.get("RANK", "0")) == 0 and not args.load_tree:
        tree.save(args.save_tree)

    for i, eta in enumerate(args.etas):
        same_init_net = None if i == 0 else f"{timestamp}-eta={args.etas[0]}"
        trial = RNaD(
            use_same_init_net_as=same_init_net,
            tree=tree,
            directory_name=f"{timestamp}-eta={eta}",
            device=device,
            wandb=False,
            eta=eta,
            bounds=[args.updates],
            delta_m=[args.delta_m],
            lr=args.lr,
            gamma_averaging=0.01,
            batch_size=PRESETS[args.preset]["batch"],
            logit_clip=2,
            net_params={"type": "MLP", "max_actions": tree.max_actions, "width": 2 ** 8},
        )
        trial.run(log_mod=10, expl_mod=1, checkpoint_mod=args.delta_m)
        logging.info("eta=%s NashConv curve: %s", eta, trial.nashconv_history)


if __name__ == "__main__":
    main()
