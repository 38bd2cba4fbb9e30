"""
The learner step as one replayable unit (learn/fused.py::LearnerStep, csrc/learner_step.cu; reference rnad.py:456,
495-526): the fused tail kernel against torch's own clip_grad_norm_ / Adam / target average, one full parameter update
against the reference's gradients (tests/golden, Adam by hand), graph replay == eager == the step-by-step path, the
graphed torch tail == the eager one bit for bit, and - on a box with two GPUs - a two-rank update with the gradient
exchange inside the tail kernel == the one-rank update of the concatenated batch.
"""
import ctypes
import os
import random
import subprocess
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import close, episodes_from_golden, mlp_from_golden, t, tree_from_golden  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def seeded_tree(ragged=False, depth=3):
    from environment.tree import Tree

    np.random.seed(5)
    random.seed(5)
    torch.manual_seed(5)
    kw = dict(max_actions=3, max_transitions=2, depth_bound=depth)
    if ragged:
        kw.update(transition_threshold=0.3, depth_bound_lambda=lambda n: n.depth_bound - 1 - 2 * (random.random() < 0.5))
    tree = Tree(**kw)
    tree.generate()
    tree.to(torch.device(DEV))
    return tree


def fresh_trial(tree, batch, name, engine, **kw):
    from learn.rnad import RNaD
    from nn.net import MLP

    dev = torch.device(DEV)
    trial = RNaD(tree=tree, device=dev, directory_name=name, batch_size=batch, eta=0.2, lr=1e-3, gamma_averaging=0.01,
                 logit_clip=2, b1_adam=kw.pop("b1", 0.0), net_params={"type": "MLP", "max_actions": 3, "width": 256}, **kw)
    torch.manual_seed(21)
    trial.net = MLP(3, 256, device=dev)
    trial.net.train()
    trial.net_target, trial.net_reg, trial.net_reg_ = (MLP(3, 256, device=dev) for _ in range(3))
    trial.net_target.load_state_dict(trial.net.state_dict())
    trial.optimizer = torch.optim.Adam(trial.net.parameters(), lr=1e-3, betas=(float(trial.b1_adam), 0.999), eps=1e-8)
    trial.step_engine = engine
    return trial


def flat(module):
    return torch.cat([p.detach().flatten() for p in module.parameters()])


def assert_params_track(p_a, p_b, atol=3e-4, outliers=5e-3, lr=1e-3, steps=1):
    """Two paths whose gradients agree to rounding noise: every parameter within `atol` - except the few whose gradient
    IS rounding noise (hidden units that are active in a handful of rows): Adam divides by sqrt(v) and turns such a
    gradient, whatever its size, into a step of the order of lr, so there the two paths may differ by up to ~2 lr per
    step.  At most a fraction `outliers` of the parameters, none further apart than 2 lr per step."""
    d = (p_a.double() - p_b.double()).abs()
    assert float((d > atol).double().mean()) <= outliers, (int((d > atol).sum()), d.numel())
    assert float(d.max()) <= 2.0 * lr * steps + atol, float(d.max())


@pytest.mark.parametrize("beta1,world_like", [(0.0, False), (0.9, False), (0.0, True)])
def test_tail_kernel_matches_torch_clip_adam_and_average(beta1, world_like):
    """rnad_learner_tail alone, through the C ABI, three steps: g = G_0 / N_0 + G_1 / N_1, clip, Adam, target average."""
    import _b200

    L = _b200.lib()
    dev = torch.device(DEV)
    gen = torch.Generator().manual_seed(4)
    n = 10756
    params0 = torch.randn(n, generator=gen) * 0.1
    target0 = torch.randn(n, generator=gen) * 0.1
    counts = (123457, 98765) if world_like else (1000, 1)
    clip = 0.05 if world_like else 1e3                     # one case where the clip bites
    lr, b2, eps, gamma = 1e-3, 0.999, 1e-8, 0.01
    ref_p = torch.nn.Parameter(params0.clone())
    ref_t = target0.clone()
    opt = torch.optim.Adam([ref_p], lr=lr, betas=(beta1, b2), eps=eps)
    bufs = {k: v.to(dev) for k, v in dict(params=params0.clone(), target=target0.clone(), m=torch.zeros(n),
                                          v=torch.zeros(n), grad=torch.zeros(n), losses=torch.zeros(4)).items()}
    ctrl = torch.zeros(32, dtype=torch.uint8, device=dev)
    stats = torch.tensor([8, counts[0], counts[1], 0], dtype=torch.int32, device=dev)
    for step in range(3):
        g01 = torch.randn(2 * n, generator=gen) * (50.0 if step == 1 else 1.0)
        sums = torch.rand(4, generator=gen) * 100
        pg, ls = g01.to(dev), sums.to(dev)
        args = _b200.TailArgs()
        args.n_params = n
        args.player_grads, args.stats, args.loss_sums = pg.data_ptr(), stats.data_ptr(), ls.data_ptr()
        args.params, args.target_params = bufs["params"].data_ptr(), bufs["target"].data_ptr()
        args.exp_avg, args.exp_avg_sq = bufs["m"].data_ptr(), bufs["v"].data_ptr()
        args.flat_grad, args.losses, args.ctrl = bufs["grad"].data_ptr(), bufs["losses"].data_ptr(), ctrl.data_ptr()
        args.lr, args.beta1, args.beta2, args.eps = lr, beta1, b2, eps
        args.grad_clip, args.gamma_averaging, args.one_minus_gamma_averaging = clip, gamma, 1 - gamma
        args.world, args.rank = 1, 0
        L.rnad_learner_tail(ctypes.byref(args), _b200.stream())
        torch.cuda.synchronize()
        # the same step with torch (rnad.py:456, 514, 516-523)
        g = g01[:n] / counts[0] + g01[n:] / counts[1]
        ref_p.grad = g.clone()
        norm = torch.nn.utils.clip_grad_norm_([ref_p], clip)
        opt.step()
        ref_t = gamma * ref_p.detach() + (1 - gamma) * ref_t
        close(bufs["grad"].cpu(), ref_p.grad, rtol=1e-5, atol=1e-9)
        close(bufs["params"].cpu(), ref_p.detach(), rtol=0, atol=2e-7)
        close(bufs["target"].cpu(), ref_t, rtol=0, atol=2e-7)
        losses = bufs["losses"].cpu()
        close(losses[0], sums[0] / counts[0] + sums[1] / counts[1], rtol=1e-5)
        close(losses[1], -(sums[2] / counts[0] + sums[3] / counts[1]), rtol=1e-5)
        close(losses[2], norm, rtol=1e-5)
        assert losses[3] == 0
    host = _b200.StepCtrl.from_buffer_copy(ctrl.cpu().numpy().tobytes())
    assert host.seq == 3 and host.adam_step == 3.0 and host.error == 0


def test_full_parameter_update_against_the_reference_gradients():
    """tests/golden/regular_a3c2d3 (width 256): the reference's `__learn` gradients (rnad_grad.*), then clip / Adam /
    target average by hand - against ONE LearnerStep update computed from the same episodes."""
    import learn.fused as fused
    from learn.rnad import RNaD

    g = dict(np.load(os.path.join(REPO, "tests", "golden", "regular_a3c2d3.npz")))
    a, width = int(g["meta"][0]), int(g["meta"][3])
    assert width == 256
    tree = tree_from_golden(g, DEV)
    ep = episodes_from_golden(g, tree, DEV)
    eta, gamma, c_bar, rho_bar, alpha = (float(x) for x in g["scalars"])
    trial = RNaD(tree=tree, device=torch.device(DEV), directory_name="pytest_step_golden", eta=eta, lr=1e-3,
                 gamma_averaging=0.01, batch_size=ep.batch_size, vtrace_gamma=gamma, c_bar=c_bar, roh_bar=rho_bar,
                 b1_adam=0.0, net_params={"type": "MLP", "max_actions": a, "width": width})
    for attr, prefix in (("net", "learner"), ("net_target", "target"), ("net_reg", "reg"), ("net_reg_", "reg_")):
        setattr(trial, attr, mlp_from_golden(g, prefix, DEV))
    trial.optimizer = torch.optim.Adam(trial.net.parameters(), lr=1e-3, betas=(0.0, 0.999), eps=1e-8)
    names = [k for k, _ in trial.net.named_parameters()]
    p0 = torch.cat([t(g[f"learner.{k}"]).flatten() for k in names])
    t0 = torch.cat([t(g[f"target.{k}"]).flatten() for k in names])
    g_ref = torch.cat([t(g[f"rnad_grad.{k}"]).flatten() for k in names])      # after the reference's clip_grad_norm_

    step = trial._step_engine_for()
    assert isinstance(step, fused.LearnerStep)
    step.learn_from(ep, alpha)
    torch.cuda.synchronize()
    g_got = step.flat_grad.cpu()
    rel = ((g_got - g_ref).norm() / g_ref.norm()).item()
    assert rel < 2e-2, rel                                 # the learner's net passes run in tf32 on the tensor core
    losses = step.losses.cpu()
    close(losses[0], g["loss_v"], rtol=1e-2, atol=1e-3)
    close(losses[1], g["loss_nerd"], rtol=1e-2, atol=2e-3)
    close(losses[2], g_got.norm(), rtol=1e-4)

    def adam_by_hand(grad):
        p = torch.nn.Parameter(p0.clone())
        p.grad = grad.clone()
        torch.optim.Adam([p], lr=1e-3, betas=(0.0, 0.999), eps=1e-8).step()
        return p.detach(), 0.01 * p.detach() + (1 - 0.01) * t0

    p_mine, t_mine = adam_by_hand(g_got)                   # the optimizer arithmetic itself: tight
    close(step.flat["params"].cpu(), p_mine, rtol=0, atol=2e-7)
    close(step.flat["target"].cpu(), t_mine, rtol=0, atol=2e-7)
    close(flat(trial.net).cpu(), p_mine, rtol=0, atol=2e-7)   # the nn.Parameters are views of the flat buffers
    p_ref, t_ref = adam_by_hand(g_ref)                     # the whole update against the reference's gradients: the first
    moved = (p_ref - p0).abs() > 1e-6                      # Adam step is lr * g / (|g| + eps), i.e. +-lr wherever g != 0
    agree = ((step.flat["params"].cpu() - p_ref).abs() < 1e-5)
    assert moved.float().mean() > 0.5
    assert agree[moved].float().mean() > 0.995, agree[moved].float().mean()
    assert agree[~moved].float().mean() > 0.98


def test_graph_replay_equals_eager_equals_stepwise_path():
    """Same seeds, same initial nets: LearnerStep captured into a CUDA graph == the same calls issued eagerly (bit for
    bit), and both track the step-by-step path (normalised targets, torch's clip / Adam / average)."""
    tree = seeded_tree(ragged=True, depth=4)
    batch = 4096
    trials = {e: fresh_trial(tree, batch, f"pytest_step_{e}", e) for e in ("off", "eager", "graph")}
    p_init = flat(trials["off"].net).clone()
    hist = {e: [] for e in trials}
    for i in range(5):
        for e, trial in trials.items():
            torch.manual_seed(500 + i)
            ep = trial.learner_step(alpha=0.25 * i)
            hist[e].append((flat(trial.net).clone(), flat(trial.net_target).clone(), trial.last_losses.clone(),
                            ep.full("indices").clone(), ep.full("policy").clone()))
    assert trials["graph"]._step.graph is not None and trials["eager"]._step.graph is None and trials["off"]._step is None
    torch.cuda.synchronize()         # the tail kernel also writes the losses into pinned host memory
    assert torch.equal(trials["graph"].last_losses_host[:2], trials["graph"].last_losses.cpu())
    for i in range(5):
        for x, y in zip(hist["graph"][i], hist["eager"][i]):
            assert torch.equal(x, y), f"step {i}: graph replay and eager launch differ"
        p_s, t_s, l_s, idx_s, pol_s = hist["off"][i]
        p_g, t_g, l_g, idx_g, pol_g = hist["graph"][i]
        if i == 0:
            assert torch.equal(idx_s, idx_g) and torch.equal(pol_s, pol_g)      # same weights, same seed: same games
        # The two paths hand the backward kernel different numbers - output gradients already divided by N_p vs the
        # undivided ones - and the kernel rounds them to tf32 operands: parameter gradients agree to tf32 noise
        # (~1e-3 relative, more where a sum cancels), i.e. the Adam updates to a few per cent of one lr-sized step.
        moved = (p_s - p_init).norm()
        assert (p_g - p_s).norm() < 0.03 * moved, ((p_g - p_s).norm().item(), moved.item())
        assert_params_track(p_g, p_s, steps=i + 1)
        close(t_g, t_s, rtol=0, atol=3e-5 + 2e-5 * (i + 1))    # (the target net averages 1 % of the parameters in per step)
        # losses: the same games -> the same sums to rounding; once the nets differ by tf32 / fp16 noise a uniform falls
        # on the other side of a cumulative probability somewhere, the batches are different samples of 4,096 games,
        # and the losses agree to sampling noise only (the NeuRD loss is a small difference of O(1) terms: -0.01 .. -0.08)
        if torch.equal(idx_s, idx_g):
            close(l_g, l_s, rtol=3e-3, atol=1e-3)
        else:
            close(l_g, l_s, rtol=0, atol=3e-2)
    # Adam's step count lives on the device; a checkpoint sees it
    trials["graph"]._step.sync_optimizer(trials["graph"])
    assert float(trials["graph"].optimizer.state[next(trials["graph"].net.parameters())]["step"]) == 5.0


def test_on_policy_step_reuses_the_rollouts_own_net_outputs():
    """With the actor == the learner net the rollout's recorded logits and values ARE the learner's forward_batch outputs
    (rnad.py:373): bit-identical to the forward kernel's on the same rows, so the step evaluates only the three other
    trunks - and that three-trunk launch, several tiles per CTA, equals the five-trunk one on its outputs; the targets
    kernel fed with logits alone (pi / log_pi derived inside, net.py:76-80) equals the one fed with the forward's."""
    import _b200

    tree = seeded_tree(ragged=True, depth=4)
    trial = fresh_trial(tree, 24000, "pytest_step_reuse", "eager")      # 8 x 24,000 rows = 1,500 tiles on 148 CTAs
    step = trial._step_engine_for()
    reuse, full = dict(step._calls(reuse=True)), dict(step._calls(reuse=False))
    _b200.lib().rnad_step_control(step.ctrl.data_ptr(), 777, 0.5, _b200.stream())
    reuse["rollout"]()
    full["pack"]()
    full["forward"]()
    full["targets"]()
    torch.cuda.synchronize()
    valid = step.arena["indices"] != 0
    assert 0.3 < float(valid.float().mean()) < 0.95
    assert torch.equal(step.logits, step.fwd["logit"]), "rollout logits != learner forward logits"
    assert torch.equal(step.arena["values"], step.fwd["v"].squeeze(-1)), "rollout values != learner forward values"
    close(step.arena["policy"][valid], step.fwd["pi"][valid], rtol=0, atol=1e-6)      # fast softmax in the rollout heads
    ref = {k: step.fwd[k].clone() for k in ("v_target", "log_pi_reg", "log_pi_reg_")}
    d_logit, d_v, sums = step.d_logit.clone(), step.d_v.clone(), step.loss_sums.clone()
    for k in ref:
        step.fwd[k].fill_(float("nan"))
    reuse["pack"]()
    reuse["forward"]()
    reuse["targets"]()
    torch.cuda.synchronize()
    for k, want in ref.items():
        assert torch.equal(step.fwd[k], want), f"three-trunk launch: {k} differs from the five-trunk launch"
    assert torch.equal(step.d_v, d_v)
    close(step.d_logit, d_logit, rtol=1e-5, atol=1e-6)
    close(step.loss_sums, sums, rtol=1e-5, atol=1e-4)


def test_step_engine_follows_changes_of_what_it_captured():
    """The captured step bakes hyper-parameters, nets and the optimizer in.  RNaD compares a cheap fingerprint every step
    (no walk over the modules) and must build a new engine when any of it changes - and keep the old one otherwise."""
    from nn.net import MLP

    tree = seeded_tree()
    trial = fresh_trial(tree, 1024, "pytest_step_key", "graph")
    for i in range(3):
        torch.manual_seed(40 + i)
        trial.learner_step(alpha=0.5)
    first = trial._step
    assert first is not None and first.graph is not None
    trial.learner_step(alpha=0.5)
    assert trial._step is first                                  # nothing changed: the same engine, the fast path
    trial.eta = 0.3                                              # a hyper-parameter the targets kernel takes by value
    trial.learner_step(alpha=0.5)
    second = trial._step
    assert second is not first and abs(second._params.eta - 0.3) < 1e-7
    trial.optimizer.param_groups[0]["lr"] = 5e-4                 # the tail kernel's learning rate
    trial.learner_step(alpha=0.5)
    third = trial._step
    assert third is not second and abs(third._tail.lr - 5e-4) < 1e-10
    trial.net_reg = MLP(3, 256, device=torch.device(DEV))        # another regularisation net (as at the end of an eta)
    trial.learner_step(alpha=0.5)
    assert trial._step is not third
    torch.cuda.synchronize()
    assert bool(torch.isfinite(flat(trial.net)).all())


def test_logging_step_in_between_keeps_the_optimizer_state_consistent():
    """A step on the step-by-step path (as wandb logging takes) between graph steps shares params, moments and step count."""
    tree = seeded_tree()
    a = fresh_trial(tree, 2048, "pytest_step_mixed_a", "graph", b1=0.9)
    b = fresh_trial(tree, 2048, "pytest_step_mixed_b", "off", b1=0.9)
    for i in range(6):
        for trial in (a, b):
            torch.manual_seed(900 + i)
            log = {} if (i == 3 and trial is a) else None
            trial.learner_step(alpha=1.0, log=log)
            if log is not None:
                assert set(log) >= {"loss_v", "loss_nerd", "gradient_norm", "entropy", "actor_learner_kld"}
    assert_params_track(flat(a.net), flat(b.net), steps=6)      # (rounding noise between the two paths, see above)
    close(flat(a.net_target), flat(b.net_target), rtol=0, atol=3e-5 + 2e-5 * 6)
    assert float(a._step.read_ctrl().adam_step) == 6.0


def test_graphed_torch_tail_equals_eager_tail_bit_for_bit():
    """The step-by-step path's optimizer tail (clip_grad_norm_, Adam, target average: ~20 torch launches) replayed as a
    CUDA graph == the same ops launched eagerly, bit for bit; and == the default (non-capturable) Adam to fp32 rounding."""
    import learn.rnad as rnad_mod

    tree = seeded_tree()
    graphed = fresh_trial(tree, 2048, "pytest_tail_graph", "off", b1=0.9)
    eager = fresh_trial(tree, 2048, "pytest_tail_eager", "off", b1=0.9)
    plain = fresh_trial(tree, 2048, "pytest_tail_plain", "off", b1=0.9)
    plain.graph_optimizer_tail = False

    class NeverCaptures(rnad_mod._GraphedTail):      # same capturable Adam, every step launched eagerly
        def step(self, trial):
            trial._eager_tail(clip=True)

    for i in range(5):
        for trial in (graphed, eager, plain):
            torch.manual_seed(700 + i)
            trial.learner_step(alpha=0.5)
            if trial is eager and not isinstance(trial._tail, NeverCaptures):
                # learner_step has just built a _GraphedTail (first step: eager warm-up); swap in the never-capturing one
                trial._tail.__class__ = NeverCaptures
    assert graphed._tail.graph is not None and eager._tail.graph is None and plain._tail is None
    assert torch.equal(flat(graphed.net), flat(eager.net)) and torch.equal(flat(graphed.net_target), flat(eager.net_target))
    close(flat(graphed.net), flat(plain.net), rtol=0, atol=5e-6)
    close(flat(graphed.net_target), flat(plain.net_target), rtol=0, atol=5e-6)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run under gpurun --gpus 2)")
@pytest.mark.timeout(600)
def test_two_rank_step_equals_single_rank(tmp_path):
    """Two ranks, each with its own half of the batch, exchange (G_0 | G_1 | N_0, N_1 | loss sums) inside the tail kernel
    over peer memory; the update must equal ONE rank's update from the concatenated batch - on a ragged tree, where the
    ranks' step counts differ - and both ranks must hold bit-identical parameters."""
    sys.path.insert(0, os.path.join(REPO, "tests", "tools"))
    import dp_step_worker as worker

    batch, steps = 3000, 4
    env = dict(os.environ, NCCL_DEBUG="WARN")
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29611",
                           os.path.join(REPO, "tests", "tools", "dp_step_worker.py"), str(tmp_path), str(batch), str(steps)],
                          env=env, capture_output=True, text=True, timeout=500)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    r0, r1 = (torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(2))
    for i in range(steps):
        assert torch.equal(r0[f"params.{i}"], r1[f"params.{i}"]), f"the ranks' parameters drifted apart at step {i}"
        assert torch.equal(r0[f"target.{i}"], r1[f"target.{i}"]) and torch.equal(r0[f"losses.{i}"], r1[f"losses.{i}"])
    assert torch.equal(r0["flat_grad"], r1["flat_grad"]) and r0["seed"] == r1["seed"]
    n0 = [(r["ep.indices"] != 0).sum().item() for r in (r0, r1)]
    assert n0[0] != n0[1], "the ranks should have played different games"

    trial = worker.build_trial(torch.device(DEV), 2 * batch, "pytest_dp_single")
    step = trial._step_engine_for()

    class Both:
        pass

    both = Both()
    for key in ("indices", "turns", "observations", "policy", "actions", "rewards", "values", "masks"):
        setattr(both, key, torch.cat([r0["ep." + key], r1["ep." + key]], dim=1).to(DEV))
    step.learn_from(both, 0.5)
    torch.cuda.synchronize()
    scale = r0["flat_grad"].abs().max().item()
    close(step.flat_grad.cpu(), r0["flat_grad"], rtol=1e-4, atol=2e-6 * scale)
    close(step.losses.cpu()[:3], r0["losses.0"][:3], rtol=1e-5, atol=1e-6)
    close(step.flat["params"].cpu(), r0["params.0"], rtol=0, atol=1e-6)
    close(step.flat["target"].cpu(), r0["target.0"], rtol=0, atol=1e-6)
