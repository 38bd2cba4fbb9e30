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


class PeerExchange:
    """
    The exchange buffers of `rnad_learner_tail` (csrc/learner_step.cu): one cudaMalloc'ed buffer per rank, mapped into
    every other rank's process through CUDA IPC, so that the tail kernel pushes its gradient row straight into its
    peers' memory over NVLink and no collective library sits on the step's critical path.  torch.distributed only
    carries the 64-byte handles once, at construction.  `pointers[r]` is rank r's buffer as seen from this process.
    """

    def __init__(self, n_params: int, device: torch.device):
        import ctypes

        import _b200

        d = group()
        assert d is not None, "PeerExchange needs an initialised multi-rank process group"
        L = _b200.lib()
        self.world, self.rank, self.device = d.get_world_size(), d.get_rank(), device
        if self.world > _b200.MAX_PEERS:
            raise _b200.RnadError(f"at most {_b200.MAX_PEERS} ranks share gradients through peer memory")
        self._lib, self._local, self._opened = L, None, []
        n_bytes = int(L.rnad_xchg_bytes(n_params, self.world))
        error = None
        handle = ctypes.create_string_buffer(_b200.IPC_HANDLE_BYTES)
        with torch.cuda.device(device):
            try:
                local = ctypes.c_void_p()
                L.rnad_xchg_create(n_bytes, ctypes.byref(local), handle)
                self._local = local
            except _b200.RnadError as exc:
                error = str(exc)
            replies = all_gather_object((self.rank, handle.raw, error))
            self.pointers = [None] * self.world
            if not any(e for _, _, e in replies):
                for r, raw, _ in replies:
                    if r == self.rank:
                        self.pointers[r] = self._local.value
                        continue
                    try:
                        peer = ctypes.c_void_p()
                        L.rnad_xchg_open(raw, ctypes.byref(peer))
                        self._opened.append(peer)
                        self.pointers[r] = peer.value
                    except _b200.RnadError as exc:
                        error = str(exc)
            errors = [e for e in all_gather_object(error) if e]
        if errors:
            self.close()
            raise _b200.RnadError("peer-memory gradient exchange unavailable: " + errors[0])

    def close(self):
        for peer in self._opened:
            try:
                self._lib.rnad_xchg_close(peer)
            except Exception:
                pass
        self._opened = []
        if self._local is not None:
            try:
                self._lib.rnad_xchg_destroy(self._local)
            except Exception:
                pass
            self._local = None
