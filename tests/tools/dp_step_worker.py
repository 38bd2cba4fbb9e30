"""
Worker of tests/test_gpu_learner_step.py::test_two_rank_step_equals_single_rank (launched under torchrun, one rank per
GPU, NCCL for the plumbing): every rank plays its own games of the SAME seeded ragged tree with the SAME nets and runs
LearnerStep updates - the gradient exchange happens inside rnad_learner_tail over CUDA-IPC peer memory - and saves
what the test needs to redo the first step on one GPU from the concatenated batch.
"""
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
for p in (os.path.join(REPO, "r-nad_b200"), REPO):
    sys.path.insert(0, p)


def build_trial(dev, batch, name):
    from environment.tree import Tree
    from learn.rnad import RNaD
    from nn.net import MLP

    np.random.seed(3)
    random.seed(3)
    torch.manual_seed(3)
    tree = Tree(max_actions=3, max_transitions=2, depth_bound=4, transition_threshold=0.3,
                depth_bound_lambda=lambda t: t.depth_bound - 1 - 2 * (random.random() < 0.5))
    tree.generate()
    tree.to(dev)
    trial = RNaD(tree=tree, device=dev, directory_name=name, batch_size=batch, eta=0.2, lr=1e-3, gamma_averaging=0.01,
                 logit_clip=2, b1_adam=0.0, net_params={"type": "MLP", "max_actions": 3, "width": 256})
    torch.manual_seed(11)
    trial.net = MLP(3, 256, device=dev)
    trial.net.train()
    trial.net_target, trial.net_reg, trial.net_reg_ = (MLP(3, 256, device=dev) for _ in range(3))
    trial.net_target.load_state_dict(trial.net.state_dict())      # (reg nets keep their own random weights)
    trial.optimizer = torch.optim.Adam(trial.net.parameters(), lr=1e-3, betas=(0.0, 0.999), eps=1e-8)
    return trial


def main():
    out_dir, batch, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import learn.fused as fused

    trial = build_trial(dev, batch, f"pytest_dp_rank{rank}")
    step = trial._step_engine_for()
    assert isinstance(step, fused.LearnerStep) and step.exchange is not None and step.exchange.world == dist.get_world_size()
    rec = {}
    for i in range(steps):
        torch.manual_seed(100 + i)                 # the same rollout seed on every rank; the game ids differ (rank << 40)
        ep = trial.learner_step(alpha=0.5)
        if i == 0:
            for key in ("indices", "turns", "observations", "policy", "actions", "rewards", "values", "masks"):
                rec["ep." + key] = ep.full(key).cpu().clone()
            rec["flat_grad"] = step.flat_grad.cpu().clone()
            rec["seed"] = ep.states.seed
        rec[f"params.{i}"] = step.flat["params"].cpu().clone()
        rec[f"target.{i}"] = step.flat["target"].cpu().clone()
        rec[f"losses.{i}"] = step.losses.cpu().clone()
    step.check()
    torch.save(rec, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    step.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
