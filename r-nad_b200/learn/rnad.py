"""
The R-NaD training loop - API mirror of the reference `learn/rnad.py` (`RNaD`
:18-547): same constructor keywords, `run(max_updates, checkpoint_mod, expl_mod,
log_mod)`, public state (`m, n, total_steps, net, net_target, net_reg, net_reg_,
optimizer, directory`), the same `saved_runs/<name>/params` and
`saved_runs/<name>/<m>/<n>` checkpoint files and auto-resume rules, so the
reference's `main.py` drives it unchanged.

One learner step (`while self.n < delta_m`, rnad.py:495-526) here is
    rollout      Episodes.generate  -> one fused persistent kernel (K2)
    forward      4 x forward_batch  -> one (T*B)-row GEMM per layer and net
    targets      vtrace.learner_targets -> one fused kernel (K3): reward transform,
                 process_policy, both players' v-trace, NeuRD force, both losses and
                 d loss / d logit, d loss / d v
    backward     torch.autograd.backward([logit, v], [d_logit, d_v]) - the same
                 parameter gradients as the reference's loss.backward()
    all-reduce   (only when torch.distributed is initialised) ONE NCCL all-reduce of
                 the flat gradient between backward and clip_grad_norm_, preceded by a
                 2-int all-reduce of the normaliser counts so the data-parallel
                 gradient is exact on ragged trees too
    Adam, target-net EMA, regularisation-net rotation as in the reference.
"""

import logging
import os
import time
from typing import Dict

import torch
import torch.nn as nn

import environment.episode as episode
import environment.tree as tree
import learn.dp as dp
import learn.fused as fused
import learn.vtrace as vtrace
import nn.net as net
import util.metric as metric


class _GraphedTail:
    """
    The optimizer tail of one learner step (clip_grad_norm_, Adam, target-net average) as a CUDA graph.
    The first step after (re)building runs eagerly - it creates Adam's state - the second one captures, and from
    then on a step is one graph launch.  Anything the graph bakes in is part of the key: the optimizer object and
    its hyper-parameters, the parameter / gradient / target storages, the clip and averaging constants.
    """

    def __init__(self, trial):
        self.key = self._key(trial)
        self.graph = None
        self.warm = False
        for group in trial.optimizer.param_groups:      # step counts live on the device, so that step() never syncs
            if not group.get("capturable", False):
                group["capturable"] = True
                for p in group["params"]:
                    st = trial.optimizer.state.get(p)
                    if st and "step" in st:
                        st["step"] = torch.as_tensor(st["step"], dtype=torch.float32).to(p.device)

    @staticmethod
    def _key(trial):
        params = list(trial.net.parameters())
        return (id(trial.optimizer), id(trial.net), id(trial.net_target),
                tuple(p.data_ptr() for p in params), tuple(p.grad.data_ptr() if p.grad is not None else 0 for p in params),
                tuple(p.data_ptr() for p in trial.net_target.parameters()),
                tuple((g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]) for g in trial.optimizer.param_groups),
                float(trial.grad_clip), float(trial.gamma_averaging), torch.cuda.current_device())

    def matches(self, trial):
        return self.key == self._key(trial)

    def step(self, trial):
        if self.graph is not None:
            self.graph.replay()
        elif not self.warm:
            trial._eager_tail(clip=True)
            self.warm = True
        else:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                trial._eager_tail(clip=True)
            self.graph = graph
            graph.replay()


class RNaD:
    def __init__(
        self,
        tree: tree.Tree,
        device=torch.device("cuda"),
        directory_name=None,
        batch_size=3 * 2 ** 8,
        eta=0.2,
        bounds=[100, 165, 200],
        delta_m=[10_000, 100_000, 35_000],
        lr=5 * 10 ** -5,
        logit_clip=2,
        neurd_clip=10 ** 3,
        grad_clip=10 ** 3,
        b1_adam=0,
        b2_adam=0.999,
        epsilon_adam=10 ** -8,
        gamma_averaging=0.001,
        roh_bar=1,
        c_bar=1,
        epsilon_threshold=0.03,
        n_discrete=32,
        n_batches_per_buffer=1,
        buffer_mod=1,
        net_params=None,
        vtrace_gamma=1,
        value_loss_weight=1,
        neurd_loss_weight=1,
        wandb=False,
        use_same_init_net_as=False,
    ):
        """Hyper-parameters as in the reference (defaults from arXiv:2206.15378; see main.py for small-scale values)."""
        self.tree = tree
        self.tree_hash = 0
        self.device = device

        self.eta = eta
        self.bounds = bounds
        self.delta_m = delta_m
        self.n_batches_per_buffer = n_batches_per_buffer
        self.buffer_mod = buffer_mod
        self.lr = lr
        self.beta = logit_clip
        self.neurd_clip = neurd_clip
        self.grad_clip = grad_clip
        self.b1_adam = b1_adam
        self.b2_adam = b2_adam
        self.epsilon_adam = epsilon_adam
        self.gamma_averaging = gamma_averaging
        self.roh_bar = roh_bar
        self.c_bar = c_bar
        self.batch_size = batch_size
        self.epsilon_threshold = epsilon_threshold
        self.n_discrete = n_discrete
        self.vtrace_gamma = vtrace_gamma
        self.neurd_weight = neurd_loss_weight
        self.value_weight = value_loss_weight
        self.wandb = wandb

        if directory_name is None:
            directory_name = str(int(time.perf_counter()))
        self.directory_name = directory_name

        if net_params is None:
            net_params = {"type": "MLP", "max_actions": self.tree.max_actions, "width": 2 ** 8}
        self.net_params = net_params

        self.saved_keys = [key for key in self.__dict__.keys() if key != "tree"]
        # only the members above are written to / restored from the 'params' file

        self._saved_runs_dir = os.path.join(os.path.dirname(os.path.realpath(__file__)), "..", "saved_runs")
        os.makedirs(self._saved_runs_dir, exist_ok=True)
        self.directory = os.path.join(self._saved_runs_dir, directory_name)
        self.use_same_init_net_as = use_same_init_net_as

        self.m = 0
        self.n = 0
        self.total_steps = 0
        self.net: nn.Module = None
        self.net_target: nn.Module = None
        self.net_reg: nn.Module = None
        self.net_reg_: nn.Module = None
        self.optimizer = None
        self.last_losses = None      # device tensor [loss_v, loss_nerd] of the latest step
        self.last_losses_host = None # (step engine) pinned host tensor [loss_v, loss_nerd, |grad|, error word], current after a synchronize
        self.nashconv_history = []   # (total_steps, NashConv of the target net)
        self.learner_engine = None   # None = auto: "fused" tensor-core kernels where supported, else "torch"
        self._fused = None
        self.graph_optimizer_tail = os.environ.get("RNAD_GRAPH_TAIL", "1") != "0"   # see _GraphedTail
        self._tail = None
        self._defer_clip = False
        # the whole learner step as one CUDA graph (learn/fused.py::LearnerStep); "0" keeps the step-by-step path,
        # "eager" runs the same five kernels calls without capturing them
        self.step_engine = os.environ.get("RNAD_STEP_ENGINE", "graph")
        self._step = None

    # ------------------------------------------------------------------ nets

    def __new_net(self) -> nn.Module:
        kinds = {"ConvNet": net.ConvNet, "MLP": net.MLP}
        cls = kinds[self.net_params["type"]]
        kwargs = {k: v for k, v in self.net_params.items() if k != "type"}
        kwargs["device"] = self.device
        new_net = cls(**kwargs)
        new_net.eval()
        return new_net

    def __new_optimizer(self):
        # the reference passes betas=[0, 0.999] (int, float), which torch >= 2.6 rejects (rnad.py:232-237)
        return torch.optim.Adam(self.net.parameters(), lr=self.lr, betas=(float(self.b1_adam), float(self.b2_adam)),
                                eps=self.epsilon_adam)

    def _is_writer(self):
        return dp.rank() == 0

    # ------------------------------------------------------- init / checkpoints

    def __initialize(self):
        """
        New run: four identical nets + Adam, checkpoint (0, 0).  Existing run: restore params and the latest checkpoint.
        Under data parallelism rank 0 alone looks at the run directory; the decision (and the checkpoint to resume
        from) is broadcast, so that every rank takes the same branch and issues the same collectives, and the other
        ranks read the files only after rank 0 has written them.
        """
        logging.info("Initializing R-NaD run: {}".format(self.directory_name))
        if dp.group() is not None:
            hashes = dp.all_gather_object(self.tree.hash)
            assert all(h == hashes[0] for h in hashes), f"the ranks hold different trees (hashes {hashes})"
        saved_updates, resume_at = [], None
        if self._is_writer():
            os.makedirs(self.directory, exist_ok=True)
            saved_updates = [int(os.path.relpath(f.path, self.directory)) for f in os.scandir(self.directory) if f.is_dir()]
            if saved_updates:
                m = max(saved_updates)
                last_update = os.path.join(self.directory, str(m))
                checkpoints = [int(os.path.relpath(f.path, last_update)) for f in os.scandir(last_update) if not f.is_dir()]
                if checkpoints:
                    resume_at = (m, max(checkpoints))
                elif m == 0:
                    saved_updates = []      # an interrupted first start left '0/' without a checkpoint: start over
                else:
                    raise FileNotFoundError(f"{last_update} holds no checkpoint")
        saved_updates, resume_at = dp.broadcast_object((saved_updates, resume_at))
        if not saved_updates:
            self.tree_hash = self.tree.hash
            if self._is_writer():
                torch.save({key: self.__dict__[key] for key in self.saved_keys}, os.path.join(self.directory, "params"))
                os.makedirs(os.path.join(self.directory, "0"), exist_ok=True)
            self.net = self.__new_net()
            if self.use_same_init_net_as:
                checkpoint = torch.load(os.path.join(self._saved_runs_dir, self.use_same_init_net_as, "0", "0"),
                                        map_location=self.device)
                self.net.load_state_dict(checkpoint["net"])
                logging.info("Loading init net from {}".format(self.use_same_init_net_as))
            dp.broadcast_parameters(self.net)   # every rank starts from rank 0's weights
            self.net.train()
            self.net_target = self.__new_net()
            self.net_target.load_state_dict(self.net.state_dict())
            self.net_reg = self.__new_net()
            self.net_reg.load_state_dict(self.net.state_dict())
            self.net_reg_ = self.__new_net()
            self.net_reg_.load_state_dict(self.net.state_dict())
            self.optimizer = self.__new_optimizer()
            self.m = 0
            self.n = 0
            self.__save_checkpoint()
            dp.barrier()                        # checkpoint 0/0 exists before anybody may want it (use_same_init_net_as)
        else:
            params_dict = torch.load(os.path.join(self.directory, "params"), weights_only=False)
            for key, value in params_dict.items():
                if key == "directory_name":
                    params_dict[key] = self.directory_name
                    continue
                if key == "device":
                    continue
                if torch.is_tensor(value):
                    params_dict[key] = params_dict[key].to(self.device)
                if key == "tree_hash":
                    assert params_dict["tree_hash"] == self.tree.hash
                self.__dict__[key] = value
            if self._is_writer():
                torch.save(params_dict, os.path.join(self.directory, "params"))
            self.m, self.n = resume_at
            self.__load_checkpoint(self.m, self.n)
            for module in (self.net, self.net_target, self.net_reg, self.net_reg_):
                dp.broadcast_parameters(module)   # same collectives on both branches; a no-op in content

        if self.wandb:
            import wandb

            wandb.init(resume=bool(saved_updates), project="RNaD",
                       config={key: self.__dict__[key] for key in self.saved_keys})
            wandb.run.name = self.directory_name

    def __load_checkpoint(self, m, n):
        saved_dict = torch.load(os.path.join(self.directory, str(m), str(n)), map_location=self.device,
                                weights_only=False)
        self.total_steps = saved_dict["total_steps"]
        self.net_params = saved_dict["net_params"]
        for name in ("net", "net_target", "net_reg", "net_reg_"):
            new = self.__new_net()
            new.load_state_dict(saved_dict[name])
            setattr(self, name, new)
        self.optimizer = self.__new_optimizer()
        self.optimizer.load_state_dict(saved_dict["optimizer"])

    def __save_checkpoint(self):
        if not self._is_writer():
            return
        if self._step is not None:
            self._step.sync_optimizer(self)      # Adam's step count lives on the device while the step engine runs
        saved_dict = {
            "total_steps": self.total_steps,
            "net_params": self.net_params,
            "net": self.net.state_dict(),
            "net_target": self.net_target.state_dict(),
            "net_reg": self.net_reg.state_dict(),
            "net_reg_": self.net_reg_.state_dict(),
            "optimizer": self.optimizer.state_dict(),
        }
        os.makedirs(os.path.join(self.directory, str(self.m)), exist_ok=True)
        torch.save(saved_dict, os.path.join(self.directory, str(self.m), str(self.n)))

    def __get_update_info(self):
        """(may continue?, delta_m of the schedule segment `m` falls in) (rnad.py:321-332)."""
        bounding = [i for i, bound in enumerate(self.bounds) if bound > self.m]
        if not bounding:
            return False, 0
        return True, self.delta_m[min(bounding)]

    # ---------------------------------------------------------------- metric

    def __nashconv(self) -> float:
        """NashConv of the target net at the root; logs the per-depth means (rnad.py:334-351)."""
        logging.info("NashConv at m: {}, n: {}, step {}".format(self.m, self.n, self.total_steps))
        nashconv_data = metric.NashConvData(self.tree)
        nashconv_data.get_nashconv_from_net(self.tree, self.net_target)
        for depth, nashconv in nashconv_data.mean_nashconv_by_depth().items():
            logging.info("depth:{}, nash_conv:{}".format(depth, nashconv))
        return (nashconv_data.row_best[1] + nashconv_data.col_best[1]).item()

    # ------------------------------------------------------------ learner step

    def __learn(self, episodes: episode.Episodes, alpha: float, log: dict = None):
        """Gradients of the learner net from a batch of trajectories (rnad.py:353-456)."""
        engine = fused.engine_for(self.net, self.learner_engine)
        pi_target = None
        if engine == "fused":
            if self._fused is None or self._fused.device != next(self.net.parameters()).device:
                self._fused = fused.FusedLearner(self.net)
            # full-length tensors: no wait for t_eff; slots past the end of a game are invalid (index 0) and masked
            observations = episodes.full("observations") if hasattr(episodes, "full") else episodes.observations
            f = self._fused.forward(observations, self.net, self.net_target, self.net_reg, self.net_reg_)
            logit, log_pi, pi, v = f["logit"], f["log_pi"], f["pi"], f["v"]
            v_target, log_pi_reg, log_pi_reg_ = f["v_target"], f["log_pi_reg"], f["log_pi_reg_"]
        else:
            logit, log_pi, pi, v = self.net.forward_batch(episodes)
            with torch.no_grad():
                _, _, pi_target, v_target = self.net_target.forward_batch(episodes)
                _, log_pi_reg, _, _ = self.net_reg.forward_batch(episodes)
                _, log_pi_reg_, _, _ = self.net_reg_.forward_batch(episodes)

        global_counts = None
        if dp.group() is not None:
            # exact data-parallel losses: normalise by the GLOBAL number of steps each player took
            global_counts = dp.all_reduce_counts(vtrace.count_played(episodes))

        out = vtrace.learner_targets(
            episodes, logit, pi, log_pi, v, v_target, log_pi_reg, log_pi_reg_,
            alpha=alpha, eta=self.eta, lambda_=1.0, c=self.c_bar, rho=self.roh_bar, gamma=self.vtrace_gamma,
            epsilon_threshold=self.epsilon_threshold, n_discrete=self.n_discrete, neurd_clip=self.neurd_clip,
            beta=self.beta, value_weight=self.value_weight, neurd_weight=self.neurd_weight,
            global_counts=global_counts)
        self.last_losses = out.losses
        self.last_losses_host = None
        if engine == "fused":
            flat = self._fused.backward(observations, self.net, out.d_logit, out.d_v)
            dp.all_reduce_flat(flat)                      # the params' .grad are views of this buffer
        else:
            torch.autograd.backward([logit, v], [out.d_logit, out.d_v.unsqueeze(-1)])
            dp.all_reduce_gradients(self.net.parameters())   # no-op unless torch.distributed is initialised

        if log is not None:
            with torch.no_grad():
                if pi_target is None:
                    _, _, pi_target, _ = self.net_target.forward_batch(episodes)
                valid = (episodes.indices != 0).to(torch.float)
                masks = episodes.masks
                # the fused engine ran on the full-length (t_max) tensors; the log is over the stored t_eff + 1 half-moves
                n_t = valid.shape[0]
                logit, pi, pi_target = logit[:n_t], pi[:n_t], pi_target[:n_t]
                total_norm = torch.sqrt(sum(p.grad.detach().pow(2).sum() for p in self.net.parameters())).item()
                logit_mean = logit.mean().item()
                uniform_policy = torch.nn.functional.normalize(masks, p=1, dim=-1)
                losses = out.losses.tolist()   # under data parallelism: this rank's share of the global loss
                log.update({
                    "loss_v": losses[0],
                    "loss_nerd": losses[1],
                    "traj_len": valid.sum(0).mean(-1).item(),
                    "gradient_norm": total_norm,
                    "logit_mean": logit_mean,
                    "logit_max": torch.max(torch.abs(logit - logit_mean)).item(),
                    "entropy": metric.kld(pi, uniform_policy, valid, legal_actions=masks),
                    "entropy_target": metric.kld(pi_target, uniform_policy, valid, legal_actions=masks),
                    "actor_learner_kld": metric.kld(pi, episodes.policy, valid, legal_actions=masks),
                })

        if not self._defer_clip:
            nn.utils.clip_grad_norm_(self.net.parameters(), self.grad_clip)

    def _eager_tail(self, clip: bool):
        """clip_grad_norm_ (rnad.py:456), Adam (514-515), target <- g * net + (1 - g) * target (516-523)."""
        if clip:
            nn.utils.clip_grad_norm_(self.net.parameters(), self.grad_clip)
        self.optimizer.step()
        with torch.no_grad():
            g = self.gamma_averaging
            src, dst = self.net.state_dict(), self.net_target.state_dict()
            float_keys = [k for k, t in dst.items() if t.is_floating_point()]
            if float_keys:
                targets = [dst[k] for k in float_keys]
                torch._foreach_mul_(targets, 1 - g)
                torch._foreach_add_(targets, [src[k] for k in float_keys], alpha=g)
            for k, t in dst.items():
                if not t.is_floating_point():   # e.g. BatchNorm's num_batches_tracked
                    t.copy_(g * src[k] + (1 - g) * t)

    def _step_engine_for(self):
        """The LearnerStep serving the current nets / optimizer / hyper-parameters, or None if it does not apply."""
        if self.step_engine in ("0", "off", "none") or self.learner_engine == "torch":
            return None
        # fast path (every step of a run): the same objects and scalars as at the last full check, and the nets'
        # first parameters still where the engine's flat buffers put them - without walking the modules
        quick = fused.LearnerStep.quick_key_of(self)
        if self._step is not None and quick == self.__dict__.get("_step_quick_key"):
            return self._step
        if os.environ.get("RNAD_LEARNER_ENGINE") == "torch" or not fused.LearnerStep.applicable(self):
            return None
        if self._step is not None and self._step.trial_key != fused.LearnerStep.key_of(self):
            self._step.sync_optimizer(self)
            self._step.close()
            self._step = None
        if self._step is None:
            try:
                self._step = fused.LearnerStep(self, use_graph=self.step_engine != "eager")
            except Exception as exc:   # e.g. CUDA IPC unavailable under data parallelism: keep the step-by-step path
                logging.warning("learner step engine unavailable (%s); using the step-by-step path", exc)
                self.step_engine = "off"
                return None
            self._tail = None
        self._step_quick_key = fused.LearnerStep.quick_key_of(self)      # (after the engine adopted the parameters)
        return self._step

    def learner_step(self, alpha: float, buffer: "episode.Buffer" = None, log: dict = None):
        """One iteration of the rnad.py:495 loop body without the schedule bookkeeping: rollout, learn, Adam, EMA."""
        if buffer is None:
            buffer = self.__dict__.setdefault("_buffer", episode.Buffer(self.n_batches_per_buffer))
        step = self._step_engine_for() if log is None else None
        if step is not None:
            # rollout -> forward -> targets -> backward -> [gradient exchange] -> clip / Adam / target average: one graph
            episodes = step.run(alpha)
            buffer.append(episodes)
            self.last_losses = step.losses[:2]
            self.last_losses_host = step.losses_host          # pinned; current after a stream synchronize
            return episodes
        if self._step is not None:
            self._step.sync_optimizer(self)          # this step runs torch's optimizer on the shared flat state
        if self.total_steps % self.buffer_mod == 0:
            episodes = episode.Episodes(self.tree, self.batch_size)
            episodes.generate(self.net)
            buffer.append(episodes)
        episodes_sample = buffer.sample(self.batch_size)
        # With the fused learner engine the gradients live in one persistent flat buffer, so everything after them -
        # clipping, Adam, the target-net average: ~20 small launches - has fixed addresses and replays as ONE CUDA graph.
        graphed = (self.graph_optimizer_tail and log is None and isinstance(self.optimizer, torch.optim.Adam)
                   and self._step is None and fused.engine_for(self.net, self.learner_engine) == "fused")
        self._defer_clip = graphed
        try:
            self.__learn(episodes_sample, alpha, log=log)
        finally:
            self._defer_clip = False
        if graphed:
            if self._tail is None or not self._tail.matches(self):
                self._tail = _GraphedTail(self)
            self._tail.step(self)
        else:
            self._eager_tail(clip=False)
            if self._step is None:
                self.optimizer.zero_grad()
        if self._step is not None:
            self._step.pull_optimizer(self)
        return episodes_sample

    def __resume(self, max_updates=10 ** 6, checkpoint_mod=1000, expl_mod=1, log_mod=20) -> None:
        """The m / n schedule (rnad.py:458-531)."""
        buffer = episode.Buffer(self.n_batches_per_buffer)
        for _ in range(max_updates):
            may_resume, delta_m = self.__get_update_info()
            if not may_resume:
                return
            logging.info("m: {}, delta_m: {}".format(self.m, delta_m))
            buffer.max_size = self.n_batches_per_buffer

            if self.m % expl_mod == 0 and self.n == 0 and self.m != 0:
                nashconv = self.__nashconv()
                self.nashconv_history.append((self.total_steps, nashconv))
                if self.wandb:
                    import wandb

                    wandb.log({"nashconv": nashconv}, step=self.total_steps)

            while self.n < delta_m:
                alpha = 1 if self.n > delta_m / 2 else self.n * 2 / delta_m
                if self.n % checkpoint_mod == 0:
                    self.__save_checkpoint()
                log = {} if (self.n % log_mod == 0 and self.wandb) else None
                self.learner_step(alpha, buffer, log=log)
                if log:
                    import wandb

                    wandb.log(log, step=self.total_steps)
                self.n += 1
                self.total_steps += 1

            self.n = 0
            self.m += 1
            self.net_reg_.load_state_dict(self.net_reg.state_dict())
            self.net_reg.load_state_dict(self.net_target.state_dict())

    def run(self, max_updates=10 ** 6, checkpoint_mod=1000, expl_mod=1, log_mod=20):
        """Starts a new run or resumes the latest checkpoint of `directory_name`."""
        self.__initialize()
        self.__resume(max_updates=max_updates, checkpoint_mod=checkpoint_mod, expl_mod=expl_mod, log_mod=log_mod)
        if self.wandb:
            import wandb

            wandb.finish()
