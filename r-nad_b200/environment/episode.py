"""
Batched self-play environment and rollout recorder - API mirror of the
reference `environment/episode.py` (`States` :18-125, `Episodes` :131-290,
`Buffer` :292-333) on top of the sm_100a kernels in csrc/ (C ABI:
include/rnad_b200.h).

What runs where
  * `States.observations` -> rnad_observe, `States.step` -> rnad_step (K1);
  * `Episodes.generate(net)` with an `nn.net.MLP` actor -> ONE launch of the
    fused persistent rollout kernel rnad_rollout (K2): gathers, both layers of
    the net, masked softmax, action and chance sampling and the direct
    (T, B, ...) trajectory writes, with a single host read (t_eff) at the end
    instead of the reference's two syncs per half-move (episode.py:96,124);
  * any other actor (e.g. ConvNet) takes the step-by-step path: the same K1
    kernels around `net.forward`, like the reference loop.

Randomness: the reference samples with the unseeded global torch generator
(`torch.multinomial`, episode.py:118, net.py:49).  Here every batch draws one
64-bit seed from torch's default generator (so `torch.manual_seed` makes runs
reproducible) and the kernels derive per-(game, half-move) uniforms from it
with Philox4x32-10 and select by inverse CDF (oracle/rnad_oracle.py pins both).
"""

import ctypes
import math
import os
import random
import time
from collections import deque

import numpy
import torch

import _b200
from environment.tree import Tree

_game_offset_rank_stride = 1 << 40   # disjoint Philox game ids per data-parallel rank


def _fresh_seed() -> int:
    """
    A new 62-bit rollout seed per batch, drawn from torch's default CPU generator: `torch.manual_seed(s)` - also the
    same `s` a second time - restarts the sequence, so a seeded run repeats its rollouts.  One scalar draw on the
    host, no device round trip.
    """
    return int(torch.empty((), dtype=torch.int64).random_(0, 1 << 62).item())


def _rank() -> int:
    import torch.distributed as dist

    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


class States:
    """A batch of games, all on the same half-move, each at some node of `tree`."""

    def __init__(self, tree: Tree, batch_size, seed=None):
        self.tree = tree
        self.batch_size = batch_size
        self._idx_tensor = None      # (B,) int32 node ids, created on first use:
        self._idx_fill = 1           # every game starts at the root, node 1 (episode.py:22)
        self._moved = False          # reference: indices is int32 until the first transition, int64 after
        self._turn = 0
        self.row_actions = None
        self.col_actions = None
        self._alive = None           # device int32 counter written by the last transition
        self._terminal = False
        self._t = 0
        self._seed = None if seed is None else int(seed)   # drawn on first use: containers (sample / collate) never roll out
        self.game_offset = _rank() * _game_offset_rank_stride

    @property
    def _idx(self) -> torch.Tensor:
        if self._idx_tensor is None:
            self._idx_tensor = torch.full((self.batch_size,), self._idx_fill, dtype=torch.int32, device=self.tree.device)
        return self._idx_tensor

    @_idx.setter
    def _idx(self, value: torch.Tensor):
        self._idx_tensor = value

    @property
    def seed(self) -> int:
        if self._seed is None:
            self._seed = _fresh_seed()
        return self._seed

    @seed.setter
    def seed(self, value):
        self._seed = int(value)

    # -- reference attributes ------------------------------------------------
    @property
    def indices(self) -> torch.Tensor:
        return self._idx.long() if self._moved else self._idx

    @indices.setter
    def indices(self, value: torch.Tensor):
        self._idx = value.to(device=self.tree.device, dtype=torch.int32).contiguous()
        self._alive = None
        self._terminal = bool(torch.all(self._idx == 0).item())

    @property
    def player_to_move(self) -> torch.Tensor:
        return torch.full((self.batch_size,), self._turn, dtype=torch.long, device=self.tree.device)

    @property
    def terminal(self) -> bool:
        """True when every game sits on the absorbing node (episode.py:124); syncs only when read."""
        if self._alive is not None:
            self._terminal = int(self._alive.item()) == 0
            self._alive = None
        return self._terminal

    # -- reference methods ---------------------------------------------------
    def observations(self) -> torch.Tensor:
        """(B, 2, A, A): expected-payoff matrix and legal mask from the mover's side (episode.py:46-68)."""
        obs, _ = self._observe(want_mask=False)
        return obs

    def _observe(self, want_mask=True):
        packed = self.tree.packed()
        a = packed.A
        with _b200.device_guard(self._idx):
            obs = torch.empty((self.batch_size, 2, a, a), dtype=torch.float32, device=self._idx.device)
            mask = torch.empty((self.batch_size, a), dtype=torch.float32, device=self._idx.device) if want_mask else None
            _b200.lib().rnad_observe(_b200.ptr(packed.ev_tab), a, packed.S, _b200.ptr(self._idx, torch.int32), self._turn,
                                     self.batch_size, _b200.ptr(obs), _b200.ptr(mask), _b200.stream())
        return obs, mask

    def observations_noisy(self) -> torch.Tensor:
        """Placeholder in the reference too (episode.py:70-82)."""
        return None

    def step(self, actions: torch.Tensor, u_chance: torch.Tensor = None) -> torch.Tensor:
        """
        Commits the mover's actions (episode.py:84-125).  Row half-move: nothing
        moves, reward 0.  Column half-move: chance is drawn, every game moves to
        its child node, and games that reach the absorbing node get their payoff.
        `u_chance` (B,) optionally injects the chance uniforms (parity tests).
        """
        dev = self._idx.device
        actions = actions.to(device=dev, dtype=torch.long).reshape(self.batch_size).contiguous()
        if self._turn == 0:
            self.row_actions = actions
            self._turn = 1
            rewards = torch.zeros((self.batch_size,), device=dev)
        else:
            self.col_actions = actions
            self._turn = 0
            packed = self.tree.packed()
            with _b200.device_guard(self._idx):
                rewards = torch.empty((self.batch_size,), dtype=torch.float32, device=dev)
                alive = torch.zeros(1, dtype=torch.int32, device=dev)
                if u_chance is not None:
                    u_chance = u_chance.to(device=dev, dtype=torch.float32).contiguous()
                _b200.lib().rnad_step(_b200.ptr(packed.tr_tab), packed.A, packed.C, _b200.ptr(self._idx, torch.int32),
                                      _b200.ptr(self.row_actions, torch.int64), _b200.ptr(self.col_actions, torch.int64),
                                      _b200.ptr(u_chance), self.seed, self._t, self.game_offset, self.batch_size,
                                      _b200.ptr(rewards), _b200.ptr(alive), _b200.stream())
            self._alive = alive
            self._moved = True
            self.row_actions = None
            self.col_actions = None
        self._t += 1
        return rewards


_scratch_cache = {}


def _rollout_scratch(dev, a, width, prec):
    """Per (device, stream, net shape, engine): the t_eff word and the kernel workspace, reused in stream order."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(),
           int(torch.cuda.current_stream(dev).cuda_stream), a, width, prec)
    hit = _scratch_cache.get(key)
    if hit is None:
        ws_bytes = int(_b200.lib().rnad_rollout_workspace_bytes(a, width, prec))
        hit = (torch.empty(1, dtype=torch.int32, device=dev),
               torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None)
        _scratch_cache[key] = hit
    return hit


_LAYOUTS = {}


class _TrajectoryArena:
    """
    One device allocation holding the eight (t_max, B, ...) trajectory tensors of a fused rollout in the order of
    `rnad_trajectory` (256-byte aligned each) plus the batch's t_eff word; tensor views are created on first use.
    """

    KEYS = ("indices", "turns", "observations", "policy", "actions", "rewards", "values", "masks")

    def __init__(self, t_max, b, a, dev):
        layout = _LAYOUTS.get((t_max, b, a))
        if layout is None:
            fields = ((torch.int64, ()), (torch.int64, ()), (torch.float32, (2, a, a)), (torch.float32, (a,)),
                      (torch.float32, (a,)), (torch.float32, ()), (torch.float32, ()), (torch.float32, (a,)))
            offsets, total = [], 0
            for dtype, tail in fields:
                n_bytes = t_max * b * math.prod(tail) * (8 if dtype == torch.int64 else 4)
                offsets.append((total, n_bytes, dtype, (t_max, b) + tail))
                total += (n_bytes + 255) // 256 * 256
            layout = _LAYOUTS[(t_max, b, a)] = (dict(zip(self.KEYS, offsets)), total)
        self.fields, self.total = layout
        self.arena = torch.empty(self.total + 256, dtype=torch.uint8, device=dev)
        # rnad_rollout's stats words: [0] = t_eff + 1, [1], [2] = valid slots of player 0 / 1 (the losses' normalisers)
        self.stats = self.arena[self.total: self.total + 16].view(torch.int32)
        self.views = {}

    def pointers(self):
        base = self.arena.data_ptr()
        return [base + self.fields[k][0] for k in self.KEYS]

    def __getitem__(self, key):
        view = self.views.get(key)
        if view is None:
            off, n_bytes, dtype, shape = self.fields[key]
            view = self.views[key] = self.arena[off: off + n_bytes].view(dtype).view(shape)
        return view


class Episodes:
    """A batch of rollout trajectories from the root; tensors are time-major (T, B, ...)."""

    TENSOR_KEYS = ("turns", "indices", "observations", "policy", "actions", "rewards", "values", "masks",
                   "q_estimates", "v_estimates")

    def __init__(self, tree: Tree, batch_size):
        self.tree: Tree = tree
        self.batch_size: int = batch_size
        self.states: States = States(tree, batch_size)
        self.finished: bool = False
        self.generation_time: float = 0
        self.estimation_time: float = 0

        self.t_eff: int = -1
        self.turns: torch.Tensor = None
        self.indices: torch.Tensor = None
        self.observations: torch.Tensor = None
        self.policy: torch.Tensor = None
        self.actions: torch.Tensor = None
        self.rewards: torch.Tensor = None
        self.values: torch.Tensor = None
        self.masks: torch.Tensor = None

        self._q_estimates: torch.Tensor = None
        self._v_estimates: torch.Tensor = None

    # After a fused rollout the trajectory length t_eff is still on the device.  Reading it is the rollout's only host
    # synchronisation, and it is deferred until somebody asks: `t_eff` and the eight (t_eff + 1, B, ...) tensors resolve
    # on first access (the kernel wrote all `t_max` half-moves; slots past t_eff sit on the absorbing node, index 0).
    # The learner's fast path never asks - it takes the full-length tensors from `full()` and lets the validity mask
    # `indices != 0` do the rest - so rollout, learner passes and optimizer are enqueued without a single host wait.
    _LAZY = ("t_eff", "indices", "turns", "observations", "policy", "actions", "rewards", "values", "masks")

    def __getattr__(self, name):
        # only reached when normal lookup fails, i.e. for a lazy attribute that has not been resolved yet
        if name in Episodes._LAZY and "_pending" in self.__dict__:
            self._resolve()
            return self.__dict__[name]
        raise AttributeError(name)

    def _resolve(self):
        pending = self.__dict__.pop("_pending", None)
        if pending is None:
            return
        full, stats = pending
        n = int(stats[0].item())
        self.__dict__["t_eff"] = n - 1
        for key in _TrajectoryArena.KEYS:
            self.__dict__[key] = full[key][:n]

    def full(self, key: str) -> torch.Tensor:
        """The (t_max, B, ...) tensor a fused rollout wrote, without waiting for t_eff; else the stored tensor."""
        pending = self.__dict__.get("_pending")
        return pending[0][key] if pending is not None else getattr(self, key)

    # The reference fills these two with zeros after every rollout (episode.py:226-227) and never reads them on the
    # training path; here they are materialised on first access.
    @property
    def q_estimates(self) -> torch.Tensor:
        if self._q_estimates is None and (self.__dict__.get("policy") is not None or "_pending" in self.__dict__):
            self._q_estimates = torch.zeros_like(self.policy)
        return self._q_estimates

    @q_estimates.setter
    def q_estimates(self, value):
        self._q_estimates = value

    @property
    def v_estimates(self) -> torch.Tensor:
        if self._v_estimates is None and (self.__dict__.get("rewards") is not None or "_pending" in self.__dict__):
            self._v_estimates = torch.zeros_like(self.rewards)
        return self._v_estimates

    @v_estimates.setter
    def v_estimates(self, value):
        self._v_estimates = value

    # ----------------------------------------------------------------- rollout

    def generate(self, net: torch.nn.Module, precision: str = None, uniforms: torch.Tensor = None):
        """
        Plays the batch to the end with `net` as the actor (episode.py:175-230).
        precision: "f16x2" (both layers of the net on tcgen05, fp16 operands: the default) | "tf32x2" (the same with tf32
        operands) | "tf32" (first
        layers on tcgen05, second on the CUDA cores) | "fp32" (CUDA cores throughout) | None =
        the net's `rollout_precision`, else $RNAD_ROLLOUT_PRECISION, else the fastest engine
        that supports the net and tree shape.  uniforms: optional (T, B, 2)
        injected action / chance uniforms (parity tests).
        """
        from nn.net import MLP

        time_start = time.perf_counter()
        if type(net) is MLP:
            self._generate_fused(net, precision, uniforms)     # the kernel reads the weights: no mode to switch
        else:
            net.eval()
            self._generate_stepwise(net)
        self.generation_time = time.perf_counter() - time_start
        self.finished = True
        if not net.training or type(net) is not MLP:
            net.train()                                        # the reference leaves the actor in training mode

    def _generate_fused(self, net, precision, uniforms):
        L = _b200.lib()
        packed = self.tree.packed()
        a, b, dev = packed.A, self.batch_size, packed.device
        if net.max_actions != a:
            raise _b200.RnadError(f"net.max_actions {net.max_actions} != tree.max_actions {a}")
        if precision is None:
            precision = getattr(net, "rollout_precision", None) or os.environ.get("RNAD_ROLLOUT_PRECISION")
        if precision is None:
            if L.rnad_rollout_tc2_supported(a, net.width, packed.C):
                precision = "f16x2"
            else:
                precision = "tf32" if L.rnad_rollout_tc_supported(a, net.width) else "fp32"
        t_max = packed.max_half_moves
        w = _b200.mlp_weights(net, dev)
        prec = _b200.PRECISIONS[precision]

        with torch.cuda.device(dev):
            # one allocation for the whole trajectory: the eight (T, B, ...) tensors are views of it, made on demand
            out = _TrajectoryArena(t_max, b, a, dev)
            traj = _b200.Trajectory(*out.pointers())
            if uniforms is not None:
                uniforms = uniforms.to(device=dev, dtype=torch.float32).contiguous()
                if uniforms.shape[0] < t_max or tuple(uniforms.shape[1:]) != (b, 2):
                    raise _b200.RnadError(f"uniforms must be ({t_max}+, {b}, 2), got {tuple(uniforms.shape)}")
                uniforms = uniforms[:t_max].contiguous()
            _, workspace = _rollout_scratch(dev, a, net.width, prec)
            stats = out.stats                                      # this batch's own stats words (the call resets them)
            L.rnad_rollout(_b200.ptr(packed.ev_tab), _b200.ptr(packed.tr_tab), a, packed.C, ctypes.byref(w), b, t_max,
                           self.states.seed, None, self.states.game_offset, _b200.ptr(uniforms), prec,
                           ctypes.byref(traj), stats.data_ptr(), _b200.ptr(workspace), _b200.stream())
        self.precision = precision
        for key in Episodes._LAZY:
            self.__dict__.pop(key, None)             # resolved on first access (see __getattr__)
        self.__dict__["_pending"] = (out, stats)
        self.__dict__["_rollout_stats"] = stats      # device int32[4]; [1], [2] are what vtrace.count_played would count
        self._q_estimates = None
        self._v_estimates = None
        self.states._idx_tensor, self.states._idx_fill = None, 0    # every game ended on the absorbing node
        self.states._moved = True
        self.states._terminal = True

    def _generate_stepwise(self, net):
        """Reference loop (episode.py:194-227) for actors the fused kernel does not cover."""
        rec = {k: [] for k in ("values", "indices", "turns", "observations", "policy", "actions", "rewards", "masks")}
        arange = torch.arange(self.batch_size, device=self.tree.device)
        while not self.states.terminal:
            rec["indices"].append(self.states.indices.clone().long())
            rec["turns"].append(self.states.player_to_move)
            observations, mask = self.states._observe()
            with torch.no_grad():
                logits, policy, value, actions = net.forward(observations)
            rewards = self.states.step(actions)
            actions_oh = torch.zeros_like(policy)
            actions_oh[arange, actions] = 1
            rec["observations"].append(observations)
            rec["values"].append(value.reshape(self.batch_size).detach().clone())
            rec["masks"].append(mask)
            rec["policy"].append(policy)
            rec["actions"].append(actions_oh)
            rec["rewards"].append(rewards)
            self.t_eff += 1
        for key, lst in rec.items():
            setattr(self, key, torch.stack(lst, dim=0))
        self.q_estimates = torch.zeros_like(self.policy)
        self.v_estimates = torch.zeros_like(self.rewards)

    # -------------------------------------------------------------- containers

    def __repr__(self):
        self._resolve()
        lines = []
        for key, value in self.__dict__.items():
            if torch.is_tensor(value) and torch.numel(value) > 20:
                value = value.shape
            lines.append(f"{key}: {value}\n")
        return "".join(lines)

    def _tensor_items(self):
        self._resolve()
        return [(k, v) for k, v in self.__dict__.items() if torch.is_tensor(v) and not k.startswith("_rollout")]

    def sample(self, batch_size):
        """A uniformly random subset (without replacement) of the batch dimension (episode.py:243-256)."""
        assert self.finished
        batch_size = min(batch_size, self.batch_size)
        selected = torch.tensor(random.sample(range(self.batch_size), batch_size), dtype=torch.long,
                                device=self.tree.device)
        result = Episodes(self.tree, batch_size)
        for key, value in self._tensor_items():
            result.__dict__[key] = torch.index_select(value, dim=1, index=selected)
        result.finished = True
        result.t_eff = self.t_eff
        return result

    @classmethod
    def collate(cls, lst: list):
        """Zero-pads every member along time to the longest and concatenates along batch (episode.py:258-290)."""
        for e in lst:
            e._resolve()
        t_eff = max(e.t_eff for e in lst)
        tree = lst[0].tree
        batch_size = sum(e.batch_size for e in lst)
        assert all(e.tree == tree for e in lst)
        assert all(e.finished for e in lst)
        result = Episodes(tree, batch_size)
        for key, _ in lst[0]._tensor_items():
            padded = []
            for e in lst:
                x = e.__dict__[key]
                extra = t_eff - e.t_eff
                if extra:
                    x = torch.cat([x, x.new_zeros((extra,) + tuple(x.shape[1:]))], dim=0)
                padded.append(x)
            result.__dict__[key] = padded[0] if len(padded) == 1 else torch.cat(padded, dim=1)
        result.finished = True
        result.t_eff = t_eff
        return result


class SelfPlay:
    """
    Repeated self-play with one actor net as a replayable unit: the trajectory lives in one static arena, the rollout
    seed in device memory, and - after one eager call - every `play()` is ONE CUDA-graph replay of

        [weights_host -> the actor's flat parameter buffer, next seed]  ->  rnad_rollout [-> returns_host]

    The seeds are a splitmix64 sequence started from one draw of torch's generator at construction and advanced on the
    device (`rnad_step_advance`); the host mirrors it, so `episodes.states.seed` names every batch's seed.

    `weights_host` (optional, pinned fp32, state_dict order, `sum(p.numel())` floats): fresh actor weights arrive from
    host memory before every batch (a learner elsewhere, a checkpoint).  `returns_host` (optional, pinned fp32, B
    floats): every game's payoff for the row player, written by the rollout kernel itself.  `play()` does not synchronise;
    the returned `Episodes` holds views of the arena, valid until the next `play()`.  (Reference: the loop around
    `Episodes.generate`, rnad.py:502-505 / episode.py:175-230.)
    """

    def __init__(self, tree: Tree, batch_size: int, net: torch.nn.Module, precision: str = None,
                 weights_host: torch.Tensor = None, returns_host: torch.Tensor = None, use_graph: bool = True):
        from nn.net import MLP

        if type(net) is not MLP:
            raise _b200.RnadError("SelfPlay serves nn.net.MLP actors (the fused rollout kernel)")
        L = _b200.lib()
        self.tree, self.net, self.batch_size = tree, net, int(batch_size)
        self.packed = packed = tree.packed()
        self.device = dev = packed.device
        if precision is None:
            if L.rnad_rollout_tc2_supported(packed.A, net.width, packed.C):
                precision = "f16x2"
            else:
                precision = "tf32" if L.rnad_rollout_tc_supported(packed.A, net.width) else "fp32"
        self.precision = precision
        self.t_max = packed.max_half_moves
        self.weights_host, self.returns_host = weights_host, returns_host
        with torch.cuda.device(dev):
            self.arena = _TrajectoryArena(self.t_max, self.batch_size, packed.A, dev)
            self._seed_state = _fresh_seed()
            host = _b200.StepCtrl()
            host.seed_state[0], host.seed_state[1] = self._seed_state & 0xFFFFFFFF, self._seed_state >> 32
            self.ctrl = torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8).to(dev)
            n = sum(p.numel() for p in net.parameters())
            self.flat = torch.empty(n, dtype=torch.float32, device=dev)
            offset = 0
            with torch.no_grad():
                for p in net.parameters():             # registration order == state_dict order
                    view = self.flat[offset: offset + p.numel()].view_as(p)
                    view.copy_(p.detach())
                    p.data = view
                    offset += p.numel()
            ws = int(L.rnad_rollout_workspace_bytes(packed.A, net.width, _b200.PRECISIONS[precision]))
            self.workspace = torch.empty(ws, dtype=torch.uint8, device=dev) if ws else None
            # returns_host is pinned, i.e. (unified addressing) device-accessible at the same address: the kernel writes
            # each game's return straight into it - 128-byte posted writes that ride under the rollout - instead of
            # a device buffer plus a copy node behind the kernel (RNAD_SELFPLAY_COPY_RETURNS=1 keeps the copy, for A/B runs)
            self.direct_returns = returns_host is not None and os.environ.get("RNAD_SELFPLAY_COPY_RETURNS") is None
            # likewise the weights: read from pinned host memory by the kernel that advances the seed
            # (RNAD_SELFPLAY_COPY_WEIGHTS=1: a copy node in front of it instead)
            self.direct_weights = os.environ.get("RNAD_SELFPLAY_COPY_WEIGHTS") is None
            self.returns = (torch.empty(self.batch_size, dtype=torch.float32, device=dev)
                            if returns_host is not None and not self.direct_returns else None)
        if weights_host is not None and (not weights_host.is_pinned() or weights_host.numel() != n):
            raise _b200.RnadError(f"weights_host must be a pinned fp32 tensor of {n} elements")
        if returns_host is not None and (not returns_host.is_pinned() or returns_host.numel() != self.batch_size):
            raise _b200.RnadError(f"returns_host must be a pinned fp32 tensor of {self.batch_size} elements")
        self._w = _b200.mlp_weights(net, dev)
        self._traj = _b200.Trajectory(*self.arena.pointers())
        if returns_host is not None:                           # the kernel sums each game's rewards itself
            self._traj.returns = returns_host.data_ptr() if self.direct_returns else self.returns.data_ptr()
        self.episodes = Episodes(tree, self.batch_size)
        self.episodes.finished = True
        self.episodes.precision = precision
        self.use_graph, self.graph, self.calls = use_graph, None, 0

    def _enqueue(self):
        L, p = _b200.lib(), self.packed
        if self.weights_host is not None and self.direct_weights:
            L.rnad_step_advance_fetch(self.ctrl.data_ptr(), self.weights_host.data_ptr(), self.flat.data_ptr(),
                                      self.flat.numel(), _b200.stream())
        else:
            if self.weights_host is not None:
                self.flat.copy_(self.weights_host, non_blocking=True)
            L.rnad_step_advance(self.ctrl.data_ptr(), _b200.stream())
        L.rnad_rollout(_b200.ptr(p.ev_tab), _b200.ptr(p.tr_tab), p.A, p.C, ctypes.byref(self._w), self.batch_size,
                       self.t_max, 0, self.ctrl.data_ptr() + _b200.StepCtrl.seed.offset, self.episodes.states.game_offset,
                       None, _b200.PRECISIONS[self.precision], ctypes.byref(self._traj), self.arena.stats.data_ptr(),
                       _b200.ptr(self.workspace), _b200.stream())
        if self.returns is not None:
            self.returns_host.copy_(self.returns, non_blocking=True)

    def _next_seed(self) -> int:
        """The host's mirror of rnad_step_advance (splitmix64)."""
        mask = 0xFFFFFFFFFFFFFFFF
        self._seed_state = (self._seed_state + 0x9E3779B97F4A7C15) & mask
        z = self._seed_state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
        return (z ^ (z >> 31)) >> 2

    def play(self) -> "Episodes":
        seed = self._next_seed()
        with torch.cuda.device(self.device):
            if self.graph is not None:
                self.graph.replay()
            elif not self.use_graph or self.calls == 0:
                self._enqueue()
            else:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._enqueue()
                self.graph = graph
                graph.replay()
        self.calls += 1
        ep = self.episodes
        ep.states.seed = seed
        for key in Episodes._LAZY:
            ep.__dict__.pop(key, None)
        ep.__dict__["_pending"] = (self.arena, self.arena.stats)
        ep.__dict__["_rollout_stats"] = self.arena.stats
        ep._q_estimates = ep._v_estimates = None
        ep.generation_time = 0.0
        return ep


class Buffer:
    """
    Replay buffer of `Episodes` batches played by older actors (episode.py:292-333).
    With the default `max_size == 1` training is on-policy; then `sample(B)` of the
    single stored batch of size B returns that batch itself (the reference returns
    a random permutation of it - every consumer is a sum over the batch, so the
    ~all-bytes-once-more permutation copy is skipped).
    """

    def __init__(self, max_size) -> None:
        self.max_size = max_size
        self.episodes_buffer = deque(maxlen=max_size)

    def sample(self, batch_size):
        n = len(self.episodes_buffer)
        if n == 1 and self.episodes_buffer[0].batch_size == batch_size and self.episodes_buffer[0].finished:
            return self.episodes_buffer[0]
        bucket_sizes = numpy.random.multinomial(batch_size, [1 / n] * n)
        assert sum(bucket_sizes) == batch_size
        return Episodes.collate([self.episodes_buffer[i].sample(int(bucket_sizes[i])) for i in range(n)])

    def append(self, episodes: Episodes):
        self.episodes_buffer.append(episodes)
        while len(self.episodes_buffer) > self.max_size:
            self.episodes_buffer.popleft()

    def clear(self):
        self.episodes_buffer.clear()
