"""
Data-parallel seam of the learner (one process per GPU, torch.distributed).

The reference has no distributed code.  Games are independent, so self-play shards
over ranks with no communication; the learner needs exactly one exchange per step:
the SUM all-reduce of the flat parameter gradient between `backward` and
`clip_grad_norm_` (rnad.py:425 / :456).  Both losses are sums over the batch divided
by per-player step counts N_p (vtrace.py:387-389, 370-374, 429), so each rank divides
its local sums by the GLOBAL N_p (one 2-int all-reduce issued before the targets
kernel) and the summed gradient equals the single-process gradient exactly - on
ragged trees too.  Backend: NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""

import torch


def group():
    """torch.distributed if a multi-rank group is initialised, else None."""
    import torch.distributed as dist

    return dist if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 else None


def rank() -> int:
    d = group()
    return d.get_rank() if d is not None else 0


def all_reduce_counts(local_counts: torch.Tensor) -> torch.Tensor:
    """Global per-player step counts (int32[2]); the input is not modified."""
    d = group()
    total = local_counts.clone()
    if d is not None:
        d.all_reduce(total)
    return total


def all_reduce_gradients(parameters) -> int:
    """Sums every parameter's .grad over ranks with ONE collective on a flat buffer; returns its element count."""
    d = group()
    grads = [p.grad for p in parameters if p.grad is not None]
    if d is None or not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    d.all_reduce(flat)
    offset = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[offset: offset + n].view_as(g))
        offset += n
    return offset


def all_reduce_flat(flat: torch.Tensor) -> None:
    """Sums an already-flat gradient buffer over ranks in place (the fused learner's params' .grad are views of it)."""
    d = group()
    if d is not None:
        d.all_reduce(flat)


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    d = group()
    if d is None:
        return
    for t in module.state_dict().values():
        d.broadcast(t, src=src)


def barrier() -> None:
    d = group()
    if d is not None:
        d.barrier()


def broadcast_object(obj, src: int = 0):
    """`obj` of rank `src` on every rank (picklable host object)."""
    d = group()
    if d is None:
        return obj
    box = [obj]
    d.broadcast_object_list(box, src=src)
    return box[0]


def all_gather_object(obj):
    d = group()
    if d is None:
        return [obj]
    out = [None] * d.get_world_size()
    d.all_gather_object(out, obj)
    return out
