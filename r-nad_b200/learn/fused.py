"""
Fused learner-side net passes (csrc/learner_mlp.cu): the four `forward_batch` calls
of reference rnad.py:373-380 as one kernel, and the parameter gradients of
rnad.py:425 (`loss.backward()`) as another, for `nn.net.MLP` nets the tensor-core
engine supports (width 256, 2 <= max_actions <= 4).  The hidden activations
(T*B x 256 per trunk) stay on the SM; first layers run on tcgen05 in tf32 with fp32
accumulation, everything after them in fp32.

`RNaD.__learn` uses this engine by default where it applies; set the environment
variable RNAD_LEARNER_ENGINE=torch (or `trial.learner_engine = "torch"`) for the
reference-style fp32 batched-GEMM + autograd path.
"""

import ctypes
import os

import torch

import _b200


def supported(net) -> bool:
    from nn.net import MLP

    if type(net) is not MLP or not next(net.parameters()).is_cuda:
        return False
    return bool(_b200.lib().rnad_learner_mlp_supported(net.max_actions, net.width))


def engine_for(net, requested=None) -> str:
    requested = requested or os.environ.get("RNAD_LEARNER_ENGINE") or "auto"
    if requested == "torch":
        return "torch"
    ok = supported(net)
    if requested == "fused" and not ok:
        raise _b200.RnadError("RNAD_LEARNER_ENGINE=fused needs CUDA nn.net.MLP nets of width 256 with 2..4 actions")
    return "fused" if ok else "torch"


class FusedLearner:
    """Buffers that persist across learner steps: kernel workspace and the flat gradient the params' .grad view."""

    def __init__(self, net):
        L = _b200.lib()
        self.a, self.width = net.max_actions, net.width
        self.device = next(net.parameters()).device
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(int(L.rnad_learner_mlp_workspace_bytes(self.a, self.width)) + 256,
                                         dtype=torch.uint8, device=self.device)
            self.n_params = int(L.rnad_learner_param_count(self.a, self.width))
            self.flat_grad = torch.zeros(self.n_params, dtype=torch.float32, device=self.device)
        assert self.workspace.data_ptr() % 256 == 0
        assert self.n_params == sum(p.numel() for p in net.parameters())

    def forward(self, observations, net, net_target, net_reg, net_reg_):
        """
        observations (T,B,2,A,A).  Returns dict: logit, pi, log_pi (T,B,A), v (T,B,1) of `net`;
        v_target (T,B,1) of `net_target`; log_pi_reg, log_pi_reg_ (T,B,A).  No autograd graph.
        """
        t, b = observations.shape[:2]
        a, n, dev = self.a, t * b, self.device
        obs = observations.detach().contiguous()
        with torch.cuda.device(dev):
            out = {k: torch.empty((t, b, a), dtype=torch.float32, device=dev)
                   for k in ("logit", "pi", "log_pi", "log_pi_reg", "log_pi_reg_")}
            out["v"] = torch.empty((t, b, 1), dtype=torch.float32, device=dev)
            out["v_target"] = torch.empty((t, b, 1), dtype=torch.float32, device=dev)
            o = _b200.LearnerFwdOut(**{k: v.data_ptr() for k, v in out.items()})
            ws = [_b200.mlp_weights(x, dev) for x in (net, net_target, net_reg, net_reg_)]
            _b200.lib().rnad_learner_forward(_b200.ptr(obs, torch.float32), n, a, *[ctypes.byref(w) for w in ws],
                                             ctypes.byref(o), _b200.ptr(self.workspace), _b200.stream())
        return out

    def backward(self, observations, net, d_logit, d_v):
        """Writes the learner's parameter gradients for (d_logit (T,B,A), d_v (T,B)) into `net`'s .grad (views of flat_grad)."""
        t, b = observations.shape[:2]
        obs = observations.detach().contiguous()
        with torch.cuda.device(self.device):
            w = _b200.mlp_weights(net, self.device)
            _b200.lib().rnad_learner_backward(_b200.ptr(obs, torch.float32), t * b, self.a, ctypes.byref(w),
                                              _b200.ptr(d_logit, torch.float32), _b200.ptr(d_v, torch.float32),
                                              _b200.ptr(self.flat_grad), _b200.ptr(self.workspace), _b200.stream())
        offset = 0
        for p in net.parameters():          # registration order == state_dict order == the kernel's flat layout
            p.grad = self.flat_grad[offset: offset + p.numel()].view_as(p)
            offset += p.numel()
        return self.flat_grad


    def backward_split(self, observations, net, d_logit, d_v):
        """
        One UNNORMALISED gradient per player (rows of even t: player 0, odd t: player 1) -> (2, P) tensor
        (`rnad_learner_backward_split`, the learner step's backward: fp16-operand engine where it exists).
        """
        t, b = observations.shape[:2]
        obs = observations.detach().contiguous()
        out = torch.empty(2, self.flat_grad.numel(), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            w = _b200.mlp_weights(net, self.device)
            _b200.lib().rnad_learner_backward_split(_b200.ptr(obs, torch.float32), t, b, self.a, ctypes.byref(w),
                                                    _b200.ptr(d_logit, torch.float32), _b200.ptr(d_v, torch.float32),
                                                    _b200.ptr(out), _b200.ptr(self.workspace), _b200.stream())
        return out


class LearnerStep:
    """
    One iteration of the reference's learner loop (rnad.py:495-526: rollout, __learn, Adam, target-net average) as a
    single replayable unit of five C-ABI calls with fixed addresses:

        rnad_rollout -> rnad_learner_forward -> rnad_learner_targets (unnormalised) -> rnad_learner_backward_split
        -> rnad_learner_tail ([peer-memory gradient exchange] + clip + Adam + target average + losses)

    The trajectory lives in one static arena, the nets' parameters and Adam's moments in flat buffers the
    `nn.Parameter`s / optimizer state are views of, the per-step scalars (rollout seed, alpha) in a 32-byte device
    control block written by `rnad_step_control`.  The first call runs the five calls eagerly, the second captures them
    into a CUDA graph, and from then on a learner update is one tiny kernel launch (the control block) plus one graph
    replay - no host synchronisation, no allocation.  Under torch.distributed the ranks exchange their gradients
    INSIDE the tail kernel through CUDA-IPC-mapped peer buffers (learn/dp.py::PeerExchange): one exchange per step
    and nothing else on NVLink.

    Serves on-policy training (`n_batches_per_buffer == 1`, `buffer_mod == 1`) of `nn.net.MLP` nets the tensor-core
    engine supports with a plain Adam optimizer; `RNaD.learner_step` keeps the step-by-step path for everything else.
    """

    def __init__(self, trial, use_graph=True):
        import learn.dp as dp
        from environment.episode import Episodes, _TrajectoryArena

        L = _b200.lib()
        net = trial.net
        self.device = dev = next(net.parameters()).device
        self.tree = trial.tree
        packed = self.tree.packed()
        self.packed = packed
        self.a, self.c, self.b, self.t = packed.A, packed.C, int(trial.batch_size), packed.max_half_moves
        self.use_graph, self.graph, self.calls, self._side = use_graph, None, 0, None
        if L.rnad_rollout_tc2_supported(self.a, net.width, packed.C):
            self.precision = "f16x2"
        else:
            self.precision = "tf32" if L.rnad_rollout_tc_supported(self.a, net.width) else "fp32"
        with torch.cuda.device(dev):
            self.n_params = n = int(L.rnad_learner_param_count(self.a, net.width))
            self.arena = _TrajectoryArena(self.t, self.b, self.a, dev)
            shape = (self.t, self.b, self.a)
            self.fwd = {k: torch.empty(shape, dtype=torch.float32, device=dev)
                        for k in ("logit", "pi", "log_pi", "log_pi_reg", "log_pi_reg_")}
            self.fwd["v"] = torch.empty((self.t, self.b, 1), dtype=torch.float32, device=dev)
            self.fwd["v_target"] = torch.empty((self.t, self.b, 1), dtype=torch.float32, device=dev)
            # the rollout's own policy-head logits: with the actor == the learner net (always, here) they and the recorded
            # values ARE the learner's forward_batch outputs (rnad.py:373), so the step does not evaluate that net again
            self.logits = torch.empty(shape, dtype=torch.float32, device=dev)
            self.reuse_actor_outputs = os.environ.get("RNAD_STEP_REUSE", "1") != "0"
            self.d_logit = torch.empty(shape, dtype=torch.float32, device=dev)
            self.d_v = torch.empty((self.t, self.b), dtype=torch.float32, device=dev)
            self.loss_sums = torch.zeros(4, dtype=torch.float32, device=dev)
            self.losses = torch.zeros(4, dtype=torch.float32, device=dev)      # loss_v, loss_nerd, |grad|, error word
            # the same four floats in pinned host memory, written by the tail kernel itself: current after a stream
            # synchronize, no copy (RNaD.last_losses_host)
            self.losses_host = torch.zeros(4, dtype=torch.float32).pin_memory()
            self.player_grads = torch.zeros(2 * n, dtype=torch.float32, device=dev)
            self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
            self.flat = {k: torch.zeros(n, dtype=torch.float32, device=dev)
                         for k in ("params", "target", "exp_avg", "exp_avg_sq")}
            self.ctrl = torch.zeros(ctypes.sizeof(_b200.StepCtrl), dtype=torch.uint8, device=dev)
            self.workspace = torch.empty(int(L.rnad_learner_mlp_workspace_bytes(self.a, net.width)) + 256,
                                         dtype=torch.uint8, device=dev)
            self.targets_ws = torch.zeros(int(L.rnad_learner_targets_workspace(self.t, self.b)), dtype=torch.uint8, device=dev)
            ws_bytes = int(L.rnad_rollout_workspace_bytes(self.a, net.width, _b200.PRECISIONS[self.precision]))
            self.rollout_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
        assert self.workspace.data_ptr() % 256 == 0 and self.ctrl.data_ptr() % 8 == 0
        self.exchange = dp.PeerExchange(n, dev) if dp.group() is not None else None
        self._adopt(trial)
        self.trial_key = self.key_of(trial)          # (after _adopt: the parameters' storages are part of the key)
        self.episodes = Episodes(self.tree, self.b)
        self.episodes.finished = True
        self.episodes.precision = self.precision
        self._build_args(trial)

    # ---- what a captured step bakes in: when any of it changes, RNaD builds a new LearnerStep
    @staticmethod
    def key_of(trial):
        opt = trial.optimizer
        group = opt.param_groups[0]
        nets = (trial.net, trial.net_target, trial.net_reg, trial.net_reg_)
        return (id(opt), tuple(id(x) for x in nets), tuple(next(x.parameters()).data_ptr() for x in nets),
                id(trial.tree), id(trial.tree.packed()), int(trial.batch_size),
                float(group["lr"]), tuple(float(b) for b in group["betas"]), float(group["eps"]),
                float(trial.grad_clip), float(trial.gamma_averaging), float(trial.eta), float(trial.c_bar),
                float(trial.roh_bar), float(trial.vtrace_gamma), float(trial.epsilon_threshold), int(trial.n_discrete),
                float(trial.neurd_clip), float(trial.beta), float(trial.value_weight), float(trial.neurd_weight))

    @staticmethod
    def quick_key_of(trial):
        """
        A cheap fingerprint of the same things (object identities, scalars, one parameter address per net) that
        does not walk the modules: RNaD compares it every step and runs `applicable` / `key_of` only when it changed.
        """
        opt = trial.optimizer
        group = opt.param_groups[0] if getattr(opt, "param_groups", None) else {}
        nets = (trial.net, trial.net_target, trial.net_reg, trial.net_reg_)
        first = tuple(getattr(getattr(x, "value_fc0", None), "weight", None) for x in nets)
        return (id(opt), type(opt), len(getattr(opt, "param_groups", ())), tuple(id(x) for x in nets),
                tuple((id(w), w.data_ptr()) if w is not None else None for w in first),
                id(trial.tree), id(getattr(trial.tree, "_packed", None)), trial.tree.index_tensor.data_ptr(),
                int(trial.batch_size), trial.step_engine, trial.learner_engine,
                os.environ.get("RNAD_LEARNER_ENGINE"), trial.n_batches_per_buffer, trial.buffer_mod,
                group.get("lr"), group.get("betas"), group.get("eps"), group.get("amsgrad"), group.get("weight_decay"),
                group.get("maximize"), group.get("capturable"), group.get("fused"), len(group.get("params", ())),
                trial.grad_clip, trial.gamma_averaging, trial.eta, trial.c_bar, trial.roh_bar, trial.vtrace_gamma,
                trial.epsilon_threshold, trial.n_discrete, trial.neurd_clip, trial.beta, trial.value_weight,
                trial.neurd_weight)

    @staticmethod
    def applicable(trial) -> bool:
        opt = trial.optimizer
        if type(opt) is not torch.optim.Adam or len(opt.param_groups) != 1:
            return False
        g = opt.param_groups[0]
        if g.get("amsgrad") or g.get("weight_decay") or g.get("maximize") or g.get("capturable") or g.get("fused"):
            return False
        nets = (trial.net, trial.net_target, trial.net_reg, trial.net_reg_)
        if not all(supported(x) for x in nets) or len({(x.max_actions, x.width) for x in nets}) != 1:
            return False
        if [id(p) for p in g["params"]] != [id(p) for p in trial.net.parameters()]:
            return False
        return trial.n_batches_per_buffer == 1 and trial.buffer_mod == 1

    def _adopt(self, trial):
        """The learner's and the target net's parameters, their .grad and Adam's moments become views of flat buffers."""
        opt = trial.optimizer
        step_count = 0.0
        offset = 0
        with torch.no_grad():
            for p, pt in zip(trial.net.parameters(), trial.net_target.parameters()):
                k = p.numel()
                sl = slice(offset, offset + k)
                self.flat["params"][sl].copy_(p.detach().reshape(-1))
                self.flat["target"][sl].copy_(pt.detach().reshape(-1))
                state = opt.state.get(p) or {}
                if "exp_avg" in state:
                    self.flat["exp_avg"][sl].copy_(state["exp_avg"].reshape(-1))
                    self.flat["exp_avg_sq"][sl].copy_(state["exp_avg_sq"].reshape(-1))
                    step_count = float(state["step"])
                p.data = self.flat["params"][sl].view_as(p)
                pt.data = self.flat["target"][sl].view_as(pt)
                p.grad = self.flat_grad[sl].view_as(p)
                opt.state[p] = {"step": torch.tensor(step_count, dtype=torch.float32),
                                "exp_avg": self.flat["exp_avg"][sl].view_as(p),
                                "exp_avg_sq": self.flat["exp_avg_sq"][sl].view_as(p)}
                offset += k
        assert offset == self.n_params
        self._write_ctrl(adam_step=step_count)

    def _write_ctrl(self, adam_step):
        host = _b200.StepCtrl()
        host.adam_step = adam_step
        self.ctrl.copy_(torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8))

    def read_ctrl(self):
        return _b200.StepCtrl.from_buffer_copy(self.ctrl.cpu().numpy().tobytes())

    def sync_optimizer(self, trial):
        """Adam's step count, kept on the device, into the torch optimizer's state (before a checkpoint or an eager step)."""
        step = torch.tensor(float(self.read_ctrl().adam_step), dtype=torch.float32)
        for p in trial.net.parameters():
            trial.optimizer.state[p]["step"] = step.clone()

    def pull_optimizer(self, trial):
        """The other direction: after torch's own optimizer.step() advanced the count."""
        ctrl = self.read_ctrl()
        ctrl.adam_step = float(trial.optimizer.state[next(trial.net.parameters())]["step"])
        self.ctrl.copy_(torch.frombuffer(bytearray(bytes(ctrl)), dtype=torch.uint8))

    def _build_args(self, trial):
        dev = self.device
        ctrl = self.ctrl.data_ptr()
        self._seed_dev = ctrl + _b200.StepCtrl.seed.offset
        self._traj = _b200.Trajectory(*self.arena.pointers())
        self._traj.logits = self.logits.data_ptr()
        self._w = [_b200.mlp_weights(x, dev) for x in (trial.net, trial.net_target, trial.net_reg, trial.net_reg_)]
        self._fwd_out = _b200.LearnerFwdOut(**{k: v.data_ptr() for k, v in self.fwd.items()})
        io = _b200.LearnerIO()
        for name, key in (("indices", "indices"), ("turns", "turns"), ("mu", "policy"), ("actions_oh", "actions"),
                          ("rewards", "rewards"), ("masks", "masks")):
            setattr(io, name, self.arena[key].data_ptr())
        for name, key in (("logit", "logit"), ("pi", "pi"), ("log_pi", "log_pi"), ("v", "v"), ("v_target_net", "v_target"),
                          ("log_pi_reg", "log_pi_reg"), ("log_pi_reg_", "log_pi_reg_")):
            setattr(io, name, self.fwd[key].data_ptr())
        io.d_logit, io.d_v = self.d_logit.data_ptr(), self.d_v.data_ptr()
        io.unnormalised, io.loss_sums = 1, self.loss_sums.data_ptr()
        self._io_full = io                 # every net output from rnad_learner_forward (given trajectories: learn_from)
        reuse = _b200.LearnerIO.from_buffer_copy(io)
        reuse.logit, reuse.v = self.logits.data_ptr(), self.arena["values"].data_ptr()
        reuse.pi = reuse.log_pi = None     # derived from the logits inside the kernel (net.py:76-80)
        self._io = reuse                   # on-policy step: the learner's own outputs are the rollout's
        self._params = _b200.LearnerParams(0.0, float(trial.eta), 1.0, float(trial.c_bar), float(trial.roh_bar),
                                           float(trial.vtrace_gamma), float(trial.epsilon_threshold), int(trial.n_discrete),
                                           float(trial.neurd_clip), float(trial.beta), float(trial.value_weight),
                                           float(trial.neurd_weight), ctrl + _b200.StepCtrl.alpha.offset)
        group = trial.optimizer.param_groups[0]
        tail = _b200.TailArgs()
        tail.n_params = self.n_params
        tail.player_grads, tail.stats, tail.loss_sums = self.player_grads.data_ptr(), self.arena.stats.data_ptr(), self.loss_sums.data_ptr()
        tail.params, tail.target_params = self.flat["params"].data_ptr(), self.flat["target"].data_ptr()
        tail.exp_avg, tail.exp_avg_sq = self.flat["exp_avg"].data_ptr(), self.flat["exp_avg_sq"].data_ptr()
        tail.flat_grad, tail.losses, tail.ctrl = self.flat_grad.data_ptr(), self.losses.data_ptr(), ctrl
        tail.losses_host = self.losses_host.data_ptr()
        tail.lr, (tail.beta1, tail.beta2), tail.eps = float(group["lr"]), (float(b) for b in group["betas"]), float(group["eps"])
        tail.grad_clip, tail.gamma_averaging = float(trial.grad_clip), float(trial.gamma_averaging)
        tail.one_minus_gamma_averaging = 1 - trial.gamma_averaging      # the reference's python-float (1 - gamma)
        tail.world, tail.rank = (self.exchange.world, self.exchange.rank) if self.exchange else (1, 0)
        if self.exchange:
            for r, pointer in enumerate(self.exchange.pointers):
                tail.xchg[r] = pointer
        self._tail = tail

    def _calls(self, reuse=None):
        """The step's C-ABI calls, in order, as (name, thunk).  `reuse`: take the learner net's outputs from the rollout."""
        L, p = _b200.lib(), self.packed
        obs = self.arena["observations"].data_ptr()
        reuse = self.reuse_actor_outputs if reuse is None else reuse
        io = self._io if reuse else self._io_full

        def rollout():
            L.rnad_rollout(_b200.ptr(p.ev_tab), _b200.ptr(p.tr_tab), self.a, self.c, ctypes.byref(self._w[0]), self.b,
                           self.t, 0, self._seed_dev, self.episodes.states.game_offset, None,
                           _b200.PRECISIONS[self.precision], ctypes.byref(self._traj), self.arena.stats.data_ptr(),
                           _b200.ptr(self.rollout_ws), _b200.stream())

        def pack():
            L.rnad_learner_pack(self.a, *[ctypes.byref(w) for w in self._w], int(reuse), _b200.ptr(self.workspace),
                                _b200.stream())

        def forward():
            L.rnad_learner_forward_prepacked(obs, self.t * self.b, self.a, *[ctypes.byref(w) for w in self._w],
                                             ctypes.byref(self._fwd_out), int(reuse), _b200.ptr(self.workspace),
                                             _b200.stream())

        def targets():
            L.rnad_learner_targets(ctypes.byref(io), ctypes.byref(self._params), self.t, self.b, self.a,
                                   _b200.ptr(self.targets_ws), _b200.stream())

        def backward():
            L.rnad_learner_backward_split_prepacked(obs, self.t, self.b, self.a, ctypes.byref(self._w[0]),
                                                    _b200.ptr(self.d_logit), _b200.ptr(self.d_v),
                                                    _b200.ptr(self.player_grads), _b200.ptr(self.workspace), _b200.stream())

        def tail():
            L.rnad_learner_tail(ctypes.byref(self._tail), _b200.stream())

        return [("pack", pack), ("rollout", rollout), ("forward", forward), ("targets", targets), ("backward", backward),
                ("tail", tail)]

    def _enqueue(self, rollout=True):
        """The step's calls in stream order - except that the weight images of the learner's net passes, which depend on
        the nets only, are packed on a side stream while the rollout runs (a fork / join the CUDA graph keeps)."""
        calls = dict(self._calls(reuse=self.reuse_actor_outputs and rollout))
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            calls["pack"]()
        if rollout:
            calls["rollout"]()
        main.wait_stream(self._side)
        for name in ("forward", "targets", "backward", "tail"):
            calls[name]()

    def profile(self, reps=20, between=None):
        """
        Device time of each of the five calls (ms, mean over `reps` eager launches, CUDA events on the launch stream;
        `between()` runs before every timed launch, e.g. an L2 flush).  Every call is a real one: `reps` learner updates
        happen.  Single rank only (the tail would wait for peers that are not stepping in lock-step).
        """
        assert self.exchange is None, "profile() is a single-rank tool"
        totals = {name: 0.0 for name, _ in self._calls()}       # ("pack" is off the critical path in a real step)
        with torch.cuda.device(self.device):
            for _ in range(reps):
                _b200.lib().rnad_step_control(self.ctrl.data_ptr(), 12345 + _, 0.5, _b200.stream())
                for name, call in self._calls():
                    if between is not None:
                        between()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    call()
                    e1.record()
                    e1.synchronize()
                    totals[name] += e0.elapsed_time(e1)
        return {name: v / reps for name, v in totals.items()}

    def run(self, alpha: float):
        """One learner update; returns the batch's `Episodes` (views of the static arena: valid until the next call)."""
        from environment.episode import Episodes, _fresh_seed

        seed = _fresh_seed()
        with torch.cuda.device(self.device):
            _b200.lib().rnad_step_control(self.ctrl.data_ptr(), seed, float(alpha), _b200.stream())
            if self.graph is not None:
                self.graph.replay()
            elif not self.use_graph or self.calls == 0:
                self._enqueue()                      # also sets the kernels' function attributes before any capture
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
        return ep

    def learn_from(self, episodes, alpha: float):
        """
        The same update from a GIVEN batch of trajectories instead of a fresh rollout (parity tests; replaying stored
        batches): `episodes` - an `Episodes` or anything with its eight (T, B, ...) tensors, T <= t_max, B == batch
        size - is copied into the arena (missing half-moves are padded with the absorbing node) and the remaining four
        calls run eagerly.
        """
        with torch.cuda.device(self.device):
            t = int(episodes.indices.shape[0])
            if t > self.t or int(episodes.indices.shape[1]) != self.b:
                raise _b200.RnadError(f"trajectory of shape {tuple(episodes.indices.shape)} does not fit ({self.t}, {self.b})")
            for key in self.arena.KEYS:
                dst = self.arena[key]
                dst.zero_()
                dst[:t].copy_(getattr(episodes, key).to(dst.dtype))
            valid = self.arena["indices"] != 0
            turns = self.arena["turns"]
            self.arena.stats.copy_(torch.stack([valid.any(1).sum(), (valid & (turns == 0)).sum(),
                                                (valid & (turns == 1)).sum(), valid.sum() * 0]).to(torch.int32))
            _b200.lib().rnad_step_control(self.ctrl.data_ptr(), 0, float(alpha), _b200.stream())
            self._enqueue(rollout=False)
        return self.losses

    def check(self):
        """Raises if a rank gave up waiting for a peer's gradients (synchronises)."""
        word = int(self.losses[3].item())
        if word:
            raise _b200.RnadError(f"gradient exchange timed out waiting for ranks with bits {word:#x}")

    def close(self):
        if self.exchange is not None:
            self.exchange.close()
            self.exchange = None
