"""
World-size-2 gloo test (CPU) of the learner's data-parallel seam (learn/dp.py):
each rank computes the loss gradients of ITS half of a ragged batch with the local
sums divided by the GLOBAL per-player step counts, the flat gradient is summed with
one all-reduce, and the result must equal the single-process gradient of the whole
batch.  The per-rank maths is the CPU oracle (test infrastructure); the code under
test is the collective plumbing RNaD.__learn uses on the GPUs with NCCL.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = os.path.join(REPO, "tests", "golden", "ragged_a3c2.npz")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _param_grads(w, ep, half, n_global, alpha, eta, gamma, c_bar, rho_bar):
    """Parameter gradients of one shard whose loss terms are normalised by `n_global` (None: by its own counts)."""
    from oracle import rnad_oracle as orc

    sub = {k: v[:, half].contiguous() for k, v in ep.items()}
    learner = {k: v.clone().requires_grad_() for k, v in w["learner"].items()}
    r = orc.learner_targets(learner, w["target"], w["reg"], w["reg_"], sub, alpha, eta, c_bar=c_bar, rho_bar=rho_bar,
                            gamma=gamma)
    d_logit, d_v = r["d_logit"].detach(), r["d_v"].detach()
    if n_global is not None:
        # oracle gradients are divided by the LOCAL counts: rescale each step by N_local / N_global of its mover
        n_local = torch.stack([hp.sum() for hp in r["has_played"]]).clamp_min(1).float()
        scale = (n_local / n_global.float().clamp_min(1))[sub["turns"]]
        d_logit = d_logit * scale.unsqueeze(-1)
        d_v = d_v * scale.unsqueeze(-1)
    torch.autograd.backward([r["logit"], r["v"]], [d_logit, d_v])
    counts = torch.stack([hp.sum() for hp in r["has_played"]]).to(torch.int32)
    return [learner[k].grad for k in sorted(learner)], counts


def _load():
    sys.path[:0] = [os.path.join(REPO, "r-nad_b200"), REPO, os.path.join(REPO, "tests")]
    from helpers import episodes_of, weights_of

    data = np.load(GOLDEN)
    g = {k: data[k] for k in data.files}
    w = {name: weights_of(g, name) for name in ("learner", "target", "reg", "reg_")}
    ep = episodes_of(g)
    eta, gamma, c_bar, rho_bar, alpha = (float(x) for x in g["scalars"])
    return w, ep, dict(alpha=alpha, eta=eta, gamma=gamma, c_bar=c_bar, rho_bar=rho_bar)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, ep, sc = _load()
    import learn.dp as dp

    assert dp.group() is not None and dp.rank() == rank
    B = ep["indices"].shape[1]
    # uneven shards of a ragged batch: the local counts differ between ranks
    bounds = [0, B // 3, B]
    half = slice(bounds[rank], bounds[rank + 1])
    _, local_counts = _param_grads(w, ep, half, None, **sc)
    global_counts = dp.all_reduce_counts(local_counts)
    assert not torch.equal(global_counts, local_counts)
    grads, _ = _param_grads(w, ep, half, global_counts, **sc)
    params = [torch.nn.Parameter(torch.zeros_like(g)) for g in grads]
    for p, g in zip(params, grads):
        p.grad = g.clone()
    n = dp.all_reduce_gradients(params)
    assert n == sum(g.numel() for g in grads)
    # parameters broadcast from rank 0
    lin = torch.nn.Linear(4, 3)
    dp.broadcast_parameters(lin)
    torch.save({"grads": [p.grad for p in params], "counts": global_counts, "lin": lin.state_dict()},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_gradient_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    w, ep, sc = _load()
    full, counts = _param_grads(w, ep, slice(None), None, **sc)
    results = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    for res in results:
        assert torch.equal(res["counts"], counts)
        for got, want in zip(res["grads"], full):
            assert torch.allclose(got, want, rtol=1e-4, atol=1e-7), (got - want).abs().max()
    for k in results[0]["lin"]:
        assert torch.equal(results[0]["lin"][k], results[1]["lin"][k])


def test_dp_is_a_noop_without_a_process_group():
    sys.path[:0] = [os.path.join(REPO, "r-nad_b200")]
    import learn.dp as dp

    assert dp.group() is None and dp.rank() == 0
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    assert dp.all_reduce_gradients([p]) == 0 and torch.equal(p.grad, torch.full((3,), 2.0))
    c = torch.tensor([3, 4], dtype=torch.int32)
    assert torch.equal(dp.all_reduce_counts(c), c)
