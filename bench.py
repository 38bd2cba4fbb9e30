#!/usr/bin/env python
"""
Benchmark of the R-NaD self-play hot path (BASELINE.json metric: self-play env
steps/s, with learner updates/s alongside).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--config cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the fused rollout (K2, which contains the K1 gathers) over
one batch of B games played from the root to the end: B * T env steps (one env step
= one game advancing one half-move).  Prints ONE JSON line.

native arm
  value     device-timed (CUDA events on the launch stream, max over ranks) kernel
            throughput with tree, weights and output buffers resident in HBM;
  e2e       the same metric through the public API (`Episodes.generate`) with the
            actor weights arriving from pinned HOST memory every step and the game
            returns read back to the host, copies inside the timed region;
  roofline  the rollout kernel against the measured HBM copy bandwidth, algorithmic
            bytes per env step from SURVEY.md section 8d (gather reads + API-faithful
            trajectory writes);
  learner   one full learner update (rollout + 4x forward_batch + K3 + backward +
            [NCCL all-reduce] + Adam + EMA) timed the same way: updates/s;
  cpu_baseline  the CPU restatement of the reference path (oracle/, "port") timed on
            this box's host cores on a bounded sample.
reference arm (--impl reference): the CPU path alone, all host threads.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(REPO, "baseline", "_ref")


def _impl_from_argv():
    for i, arg in enumerate(sys.argv):
        if arg == "--impl" and i + 1 < len(sys.argv):
            return sys.argv[i + 1]
        if arg.startswith("--impl="):
            return arg.split("=", 1)[1]
    return "native"


# The reference's packages are called environment / nn / learn / util - and so are this repository's API mirrors.
# The reference arm therefore runs in a process that never puts r-nad_b200/ on sys.path (it imports baseline/_ref/ref
# instead); everything else sees this repository's package.
REFERENCE_ARM = (_impl_from_argv() == "reference" and os.path.isfile(os.path.join(REF_ROOT, "ref", "learn", "rnad.py"))
                 and "--emit-tree" not in sys.argv)
if not REFERENCE_ARM:
    for _p in (os.path.join(REPO, "r-nad_b200"), REPO):
        if _p not in sys.path:
            sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CONFIGS = {
    # name: (depth, max_actions, max_transitions, batch per GPU)
    "cfg1": (2, 2, 1, 256),
    "cfg2": (4, 3, 2, 65536),
    "cfg2small": (3, 3, 2, 8192),
    # full regular tree of 14,900,789 nodes (7.9 GB in the reference layout, HBM-resident gathers); built on the GPU by
    # the level-synchronous generator (environment/fast_tree.py) - the reference generator would need about two days
    "cfg3": (6, 3, 3, 262144),
    # BASELINE config 4, per-GPU share (1,048,576 games over 8 GPUs): depth 8, A = 4, C = 2 thinned like main.py
    # (transition_threshold 0.3, depth_bound - 1 - 2 * (random() < 0.5)); the full regular tree would be 3.5e10 nodes
    "cfg4": (8, 4, 2, 131072),
}
FAST_TREE_CONFIGS = ("cfg3", "cfg4")


def algorithmic_bytes_per_env_step(a, c):
    """K1 reads 8A^2 + 2C + 6, K2 API-faithful trajectory writes 24 + 8A^2 + 12A."""
    return (8 * a * a + 2 * c + 6) + (24 + 8 * a * a + 12 * a)


def flops_per_env_step(a, width=256):
    return 2 * width * (4 * a * a + a + 1)


def make_tree(depth, a, c, seed=0):
    import random

    from environment.tree import Tree

    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)
    tree = Tree(max_actions=a, max_transitions=c, depth_bound=depth)
    tree.generate()
    return tree


def profiled_dram_traffic(kernel, config):
    """
    DRAM bytes per launch of `kernel` from the committed `ncu --set full` summary (profiles/), or None.
    The capture is of bench.py at cfg2 with a warm L2: most of the 69 MB trajectory is still in the 126 MB L2
    when the kernel ends, so this is a lower bound of the traffic that eventually reaches HBM.
    """
    path = os.path.join(REPO, "profiles", f"r02_{kernel}.md")
    if not os.path.exists(path):
        path = os.path.join(REPO, "profiles", f"r01_{kernel}.md")
    if config != "cfg2" or not os.path.exists(path):
        return None
    total = 0.0
    for line in open(path):
        cells = [c.strip() for c in line.split("|")]
        if len(cells) >= 4 and cells[1] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(cells[3])
            if scale is None:
                return None
            total += float(cells[2]) * scale
    return total or None


def fast_tree(config, depth, a, c, device):
    """The large configurations' trees, built level by level on `device` (environment/fast_tree.py)."""
    from environment.fast_tree import depth_jitter
    from environment.tree import Tree

    if config == "cfg4":
        tree = Tree(device=device, max_actions=a, max_transitions=c, depth_bound=depth, transition_threshold=0.3)
        tree.generate_fast(seed=0, child_spec=depth_jitter(0.5))
    else:
        tree = Tree(device=device, max_actions=a, max_transitions=c, depth_bound=depth)
        tree.generate_fast(seed=0)
    return tree


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, sm_max, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                sm_max.append(float(f[2]))
            except ValueError:
                continue
            for name, flag in zip(names, f[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(sm_max) if sm_max else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- native arm

class RolloutRunner:
    """Pre-allocated trajectory buffers + a direct C-ABI launch (no host sync) for kernel timing."""

    def __init__(self, tree, net, batch, precision):
        import ctypes

        import _b200

        self._b200, self._ctypes = _b200, ctypes
        self.L = _b200.lib()
        self.packed = tree.packed()
        p = self.packed
        dev = p.device
        self.batch, self.T, self.precision = batch, p.max_half_moves, _b200.PRECISIONS[precision]
        a, t = p.A, self.T
        self.out = {
            "indices": torch.empty((t, batch), dtype=torch.int64, device=dev),
            "turns": torch.empty((t, batch), dtype=torch.int64, device=dev),
            "observations": torch.empty((t, batch, 2, a, a), dtype=torch.float32, device=dev),
            "policy": torch.empty((t, batch, a), dtype=torch.float32, device=dev),
            "actions": torch.empty((t, batch, a), dtype=torch.float32, device=dev),
            "rewards": torch.empty((t, batch), dtype=torch.float32, device=dev),
            "values": torch.empty((t, batch), dtype=torch.float32, device=dev),
            "masks": torch.empty((t, batch, a), dtype=torch.float32, device=dev),
        }
        self.traj = _b200.Trajectory(**{k: v.data_ptr() for k, v in self.out.items()})
        self.w = _b200.MlpWeights()
        for layer in ("value_fc0", "value_fc1", "policy_fc0", "policy_fc1"):
            lin = getattr(net, layer)
            setattr(self.w, layer + "_w", lin.weight.data_ptr())
            setattr(self.w, layer + "_b", lin.bias.data_ptr())
        self.w.width = net.width
        self.stats = torch.zeros(4, dtype=torch.int32, device=dev)
        ws = int(self.L.rnad_rollout_workspace_bytes(a, net.width, self.precision))
        self.workspace = torch.empty(max(ws, 16), dtype=torch.uint8, device=dev)
        self.launches_per_step = 2 if ws else 1      # weight-image pre-kernel + rollout kernel
        self.seed = 1

    def launch(self):
        b, c = self._b200, self._ctypes
        p = self.packed
        self.seed += 1
        self.L.rnad_rollout(b.ptr(p.ev_tab), b.ptr(p.tr_tab), p.A, p.C, c.byref(self.w), self.batch, self.T,
                            self.seed, None, 0, None, self.precision, c.byref(self.traj), b.ptr(self.stats),
                            b.ptr(self.workspace), b.stream())


def timed_steps(fn, steps, warmup, flush, barrier):
    """
    W warm-up calls, then K calls each bracketed by its own pair of CUDA events on the current stream, the L2 flushed
    in between (outside the brackets).  The events exist before the loop and nothing synchronises inside it (unless
    `fn` itself does, as the end-to-end step must): the flush keeps the GPU busy while the host enqueues the next
    bracket, so a bracket holds device time only - not the host's launch latency, which grows when eight ranks share
    the box's cores.
    """
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    wall0 = time.perf_counter()
    for start, stop in zip(starts, stops):
        flush()
        start.record()
        fn()
        stop.record()
    torch.cuda.synchronize()
    total_ms = sum(start.elapsed_time(stop) for start, stop in zip(starts, stops))
    barrier()
    torch.cuda.synchronize()
    return total_ms, time.perf_counter() - wall0


def cpu_rollout_baseline(tree_tables, weights, a, batch, t_max, budget_s, threads):
    """The reference path restated on the CPU (oracle/, torch-CPU ops like the reference's own), bounded sample."""
    from oracle import rnad_oracle as orc

    torch.set_num_threads(threads)
    orc.rollout(tree_tables, weights, min(batch, 1024), t_max, seed=0)       # warm-up
    done, steps_done, t0 = 0, 0, time.perf_counter()
    while True:
        out = orc.rollout(tree_tables, weights, batch, t_max, seed=done + 1)
        steps_done += int((out["indices"] != 0).sum())
        done += 1
        elapsed = time.perf_counter() - t0
        if elapsed > budget_s or done >= 20:
            break
    return steps_done / elapsed, done, elapsed


def learner_algorithmic(a, width=256):
    """Algorithmic bytes and flops per trajectory row (one (t, b) slot) of the learner's kernels."""
    kin = 2 * a * a
    return {
        # obs read; logit, pi, log_pi, log_pi_reg, log_pi_reg_ (A floats each), v, v_target written
        "forward": {"bytes": 4 * kin + 5 * 4 * a + 8, "flops": 5 * 2 * kin * width + 2 * width * (2 + 3 * a)},
        "targets": {"bytes": 44 + 32 * a, "flops": 80},
        # obs, d_logit, d_v read; recompute of two trunks, dW1 | db1, dW2, (g W2)^T
        "backward": {"bytes": 4 * kin + 4 * a + 4, "flops": 2 * 2 * kin * width + 2 * 2 * (kin + 1) * width + 2 * 2 * width * (1 + a)},
    }


def run_native(args):
    import torch.distributed as dist

    from environment.episode import SelfPlay
    from learn.rnad import RNaD
    from nn.net import MLP

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    depth, a, c, batch = CONFIGS[args.config]
    if args.batch:
        batch = args.batch
    if args.config in FAST_TREE_CONFIGS:
        tree = fast_tree(args.config, depth, a, c, dev)
        tables = None                            # copied to the host only if the CPU baseline runs
        n_nodes = int(tree.index_tensor.shape[0])
    else:
        tree_cpu = make_tree(depth, a, c, seed=0)
        tables = {"index": tree_cpu.index_tensor.clone(), "value": tree_cpu.value_tensor.clone(),
                  "chance": tree_cpu.chance_tensor.clone(), "expected_value": tree_cpu.expected_value_tensor.clone(),
                  "legal": tree_cpu.legal_tensor.clone()}
        n_nodes = int(tree_cpu.index_tensor.shape[0])
        tree = tree_cpu
        tree.to(dev)
    torch.manual_seed(1234)
    net = MLP(a, 256, device=dev)
    weights_cpu = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    precision = args.precision
    runner = RolloutRunner(tree, net, batch, precision)
    T = runner.T

    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def flush():
        flush_buf.fill_(1.0)

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- value: kernel throughput, everything resident
    kernel_ms, _ = timed_steps(runner.launch, args.steps, args.warmup, flush, barrier)
    assert int(runner.stats[0].item()) == T
    # one env step = one VALID (t, b) slot: all of them on a regular tree, counted by the kernel itself on a ragged one
    env_steps = int(runner.stats[1].item()) + int(runner.stats[2].item())
    assert env_steps == int((runner.out["indices"] != 0).sum().item())
    assert env_steps == batch * T or args.config == "cfg4"

    # ---- the same kernel, same precision as the reference's arithmetic: the fp32 validation engine
    fp32 = None
    if args.fp32_steps > 0 and runner.L.rnad_rollout_workspace_bytes(a, 256, 0) == 0:
        runner32 = RolloutRunner(tree, net, batch, "fp32")
        ms32, _ = timed_steps(runner32.launch, args.fp32_steps, 3, flush, barrier)
        fp32 = {"ms_per_step": ms32 / args.fp32_steps, "steps": args.fp32_steps, "kernel": "rollout_fp32_kernel",
                "what": "the same rollout with the net in fp32 on the CUDA cores (the reference's arithmetic; the "
                        "validation engine that reproduces the reference's Episodes under equal uniforms)"}
        del runner32

    # ---- sustained: back-to-back rollouts for >= 1 s (no L2 flush in between: the 2 MB tables stay L2-resident as they
    # do in training, the 69 MB trajectory is rewritten every launch), one event pair around the whole run
    sustained = None
    if args.sustained_s > 0:
        n_sus = max(50, int(args.sustained_s * 1e3 / (kernel_ms / args.steps)))
        sus_sampler = ClockSampler(local_rank)
        torch.cuda.synchronize()
        barrier()
        sus_sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_sus):
            runner.launch()
        e1.record()
        e1.synchronize()
        sus_ms = e0.elapsed_time(e1)
        sustained = {"launches": n_sus, "seconds": sus_ms / 1e3, "ms_per_step": sus_ms / n_sus, "clocks": sus_sampler.stop()}

    # ---- e2e: the public API with host buffers: every step the actor's weights arrive from pinned HOST memory and the
    # per-game returns go back to pinned host memory; environment.episode.SelfPlay replays copy -> rollout [-> copy] as
    # one CUDA graph (timed region = control launch + replay + stream synchronize)
    flat_host = torch.cat([v.flatten() for v in weights_cpu.values()]).pin_memory()
    returns_host = torch.empty(batch, dtype=torch.float32).pin_memory()
    h2d_bytes = flat_host.numel() * 4
    d2h_bytes = batch * 4                       # per-game returns
    actor = MLP(a, 256, device=dev)
    actor.load_state_dict(net.state_dict())
    play = SelfPlay(tree, batch, actor, precision=precision, weights_host=flat_host, returns_host=returns_host)

    def e2e_step():
        play.play()
        torch.cuda.current_stream().synchronize()

    e2e_ms, _ = timed_steps(e2e_step, args.steps, args.warmup, flush, barrier)
    assert play.graph is not None
    assert abs(float(returns_host.abs().sum()) - float((play.arena["rewards"].abs().sum()).item())) < 1e-3 * batch

    # ---- learner: full updates through the public RNaD loop body (learn/fused.py::LearnerStep: one CUDA graph)
    trial = RNaD(tree=tree, device=dev, directory_name=f"bench_rank{rank}", batch_size=batch, eta=0.2, lr=1e-3,
                 gamma_averaging=0.01, logit_clip=2, net_params={"type": "MLP", "max_actions": a, "width": 256})
    trial.net = net
    trial.net.train()
    trial.net_target, trial.net_reg, trial.net_reg_ = (MLP(a, 256, device=dev) for _ in range(3))
    for other in (trial.net_target, trial.net_reg, trial.net_reg_):
        other.load_state_dict(net.state_dict())
    trial.optimizer = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.0, 0.999), eps=1e-8)
    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()

    def learner_step():
        trial.learner_step(alpha=0.5)
        trial.total_steps += 1
        torch.cuda.current_stream().synchronize()
        loss_host.copy_(trial.last_losses_host[:2])       # (pinned host memory the tail kernel wrote: a host-side read)

    learner_steps = max(args.learner_steps, args.steps)
    learner_ms, _ = timed_steps(learner_step, learner_steps, max(3, args.warmup), flush, barrier)
    step_engine = trial._step
    # free-running: updates enqueued back to back, one synchronisation at the end (what RNaD.run does between logs)
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(learner_steps):
        trial.learner_step(alpha=0.5)
    e1.record()
    e1.synchronize()
    free_ms = e0.elapsed_time(e1)
    if step_engine is not None:
        step_engine.check()
    per_kernel = step_engine.profile(reps=20, between=flush) if (step_engine is not None and world == 1) else None
    clocks = sampler.stop()

    # ---- max over ranks
    times = torch.tensor([kernel_ms, e2e_ms, learner_ms, free_ms], dtype=torch.float64, device=dev)
    counts = torch.tensor([env_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts)                  # ragged trees: the ranks' valid slots differ
    kernel_ms, e2e_ms, learner_ms, free_ms = (float(x) for x in times.tolist())
    total_env_steps = int(counts.item()) if world > 1 else env_steps

    result = None
    if rank == 0:
        peaks = measured_peaks()
        value = total_env_steps * args.steps / (kernel_ms / 1e3)
        per_launch_s = kernel_ms / 1e3 / args.steps
        bytes_per_launch = algorithmic_bytes_per_env_step(a, c) * env_steps
        achieved_gbs = bytes_per_launch / per_launch_s / 1e9
        tflops = flops_per_env_step(a) * env_steps / per_launch_s / 1e12
        kernel_name = {"tf32": "rollout_tc_kernel", "tf32x2": "rollout_tc2_kernel", "f16x2": "rollout_tc2_kernel",
                       "fp32": "rollout_fp32_kernel"}[precision]
        cpu = None
        if world == 1 and args.cpu_budget > 0:   # rank 0 at N = 1 only
            threads = os.cpu_count() or 1
            if tables is None:
                tables = {"index": tree.index_tensor.cpu(), "value": tree.value_tensor.cpu(),
                          "chance": tree.chance_tensor.cpu(), "expected_value": tree.expected_value_tensor.cpu(),
                          "legal": tree.legal_tensor.cpu()}
            cpu_batch = min(batch, 65536)
            cpu_value, cpu_rollouts, cpu_elapsed = cpu_rollout_baseline(tables, weights_cpu, a, cpu_batch, T,
                                                                        args.cpu_budget, threads)
            cpu = {"value": cpu_value, "unit": "env_steps/s", "cores": threads, "kind": "port",
                   "sample": f"{cpu_rollouts} rollouts of {cpu_batch} games x {T} half-moves on the same tree and net "
                             f"({cpu_elapsed:.1f} s, oracle/ restatement, torch-CPU ops, {threads} threads); the unmodified "
                             f"reference itself is timed by `bench.py --impl reference` (kind \"reference\")"}
        learner = {"updates_per_sec": learner_steps / (learner_ms / 1e3), "ms_per_update": learner_ms / learner_steps,
                   "env_steps_per_sec": total_env_steps * learner_steps / (learner_ms / 1e3),
                   "steps": learner_steps,
                   "free_running": {"ms_per_update": free_ms / learner_steps, "updates_per_sec": learner_steps / (free_ms / 1e3),
                                    "what": "the same updates enqueued back to back, one synchronisation at the end"},
                   "engine": "LearnerStep (one CUDA graph per update)" if step_engine is not None and step_engine.graph is not None
                             else "step-by-step path",
                   "exchange": (f"inside rnad_learner_tail over CUDA-IPC peer memory, {world} ranks, "
                                f"{2 * step_engine.n_params + 8} floats per rank and step, no NCCL call in the step")
                               if (step_engine is not None and step_engine.exchange is not None) else "none (one rank)",
                   "what": "RNaD.learner_step: rollout + 4x forward_batch + fused v-trace/NeuRD targets + backward"
                           + (" + gradient exchange" if world > 1 else "") + " + clip + Adam + target average, "
                           "losses read back to the host every step"}
        if per_kernel is not None:
            rows = batch * T
            alg = learner_algorithmic(a)
            kernels = {"rollout": {"ms": per_kernel["rollout"]},
                       "pack": {"ms": per_kernel["pack"], "bound": "off the critical path (side stream, under the rollout)"},
                       "tail": {"ms": per_kernel["tail"], "bound": "latency (one 8-CTA cluster, 43 KB of parameters)"}}
            for name in ("forward", "targets", "backward"):
                sec = per_kernel[name] / 1e3
                gbs, tf = alg[name]["bytes"] * rows / sec / 1e9, alg[name]["flops"] * rows / sec / 1e12
                kernels[name] = {"ms": per_kernel[name], "achieved_GBps": gbs, "frac_of_hbm": gbs / peaks["hbm_gbs"],
                                 "achieved_TFLOPs": tf, "frac_of_bf16_peak": tf / peaks["bf16_tflops"],
                                 "bound": "hbm" if name == "targets" else "tensor",
                                 "algorithmic_bytes_per_row": alg[name]["bytes"], "algorithmic_flops_per_row": alg[name]["flops"]}
            learner["roofline"] = {"kernels": kernels, "sum_of_kernels_ms": sum(v for k, v in per_kernel.items() if k != "pack"),
                                   "how": "each C-ABI call of the step timed alone with CUDA events, 20 launches, L2 flushed"}
        result = {
            "metric": "self_play_env_steps_per_sec", "value": value, "unit": "env_steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"tf32": "tf32 (tcgen05, fp32 accumulate) + fp32", "tf32x2": "tf32 (tcgen05, fp32 accumulate)",
                                          "f16x2": "fp16 operands (tcgen05 kind::f16, 11-bit significand like tf32), fp32 accumulate",
                                          "fp32": "fp32"}[precision],
            "data": "synthetic",
            "config": workload_config(args.config, depth, a, c, n_nodes, batch, T),
            "engine": {"rollout_kernel": kernel_name, "precision": precision, "env_steps_per_step": total_env_steps,
                       "partitioning": f"games sharded, {world} rank(s), no data-path collective in the rollout"},
            "clocks": clocks,
            "e2e": {"value": total_env_steps * args.steps / (e2e_ms / 1e3), "unit": "env_steps/s",
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms / args.steps,
                    "what": "environment.episode.SelfPlay.play(): net weights copied from pinned host memory, rollout, "
                            "per-game returns " + ("written by the kernel straight into" if play.direct_returns else "copied back to") +
                            " pinned host memory - one CUDA graph per batch - then a stream synchronize"},
            "gpu_launches": args.steps * runner.launches_per_step,
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved_gbs / peaks["hbm_gbs"],
                         "traffic": profiled_dram_traffic(kernel_name, args.config),
                         "kernel": kernel_name,
                         "algorithmic_bytes_per_env_step": algorithmic_bytes_per_env_step(a, c),
                         "peak_source": peaks["source"]},
            "roofline_tensor": {"achieved": tflops, "unit": "TFLOP/s (algorithmic MLP flops)",
                                "peak_bf16": peaks["bf16_tflops"], "frac_of_bf16_peak": tflops / peaks["bf16_tflops"]},
            "learner": learner,
            "cpu_baseline": cpu,
        }
        if fp32 is not None:
            fp32["value"] = total_env_steps / (fp32["ms_per_step"] / 1e3)
            fp32["unit"] = "env_steps/s"
            result["fp32"] = fp32
        if sustained is not None:
            sustained["value"] = total_env_steps / (sustained["ms_per_step"] / 1e3)
            sustained["unit"] = "env_steps/s"
            sustained["what"] = ("rollouts launched back to back for about a second (no L2 flush in between), one "
                                 "CUDA event pair around the run, clocks sampled during it")
            result["sustained"] = sustained
        print(json.dumps(result), flush=True)
    if step_engine is not None:
        step_engine.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return result


# ------------------------------------------------------------------------ reference arm

def workload_config(config, depth, a, c, n_nodes, batch, t_max):
    """The `config` object of the JSON line - the same in both arms (the driver compares them)."""
    shape = "thinned ragged" if config == "cfg4" else "regular"
    return {"workload": f"{config}: depth={depth} max_actions={a} max_transitions={c} {shape} tree ({n_nodes} nodes), "
                        f"batch={batch} games per GPU, T={t_max} half-moves, MLP width 256, self-play rollout fused "
                        f"with the policy/value net forward + one R-NaD learner update per batch",
            "l2": "256 MiB buffer written between timed steps"}


def emit_tree(args):
    """Writes the configuration's tree tables (built by this repository's generators, seed 0) to a file, for the
    reference arm: both arms then play on the same tensors."""
    depth, a, c, _ = CONFIGS[args.config]
    if args.config in FAST_TREE_CONFIGS:
        tree = fast_tree(args.config, depth, a, c, torch.device("cpu"))
    else:
        tree = make_tree(depth, a, c, seed=0)
    keys = ("index_tensor", "value_tensor", "chance_tensor", "expected_value_tensor", "legal_tensor",
            "root_value_tensor", "solution_tensor")
    torch.save({k: getattr(tree, k).cpu() for k in keys} | {"hash": tree.hash}, args.emit_tree)


def run_reference(args):
    """
    The UNMODIFIED reference (baseline/_ref/ref = a copy of /root/reference, see baseline/install_reference.py) on this
    box's host cores: `Episodes.generate` (episode.py:175-230, timed by its own `generation_time`, :192-215) and the
    learner loop body (rnad.py:495-526) through `RNaD.__resume`.  Harness-side only: the stand-in `pygambit` on
    sys.path (the real one is absent), `b1_adam=0.0` (the reference's int default breaks Adam on torch 2.11), the tree
    tensors handed over from a file so that both arms play the same tree.
    """
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    if not REFERENCE_ARM:
        return run_reference_port(args)
    import tempfile

    depth, a, c, batch = CONFIGS[args.config]
    if args.batch:
        batch = args.batch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    with tempfile.TemporaryDirectory(prefix="rnad_ref_tree_") as tmp:
        tree_file = os.path.join(tmp, "tree.pt")
        subprocess.run([sys.executable, os.path.abspath(__file__), "--emit-tree", tree_file, "--config", args.config],
                       check=True)
        tables = torch.load(tree_file)
    sys.path.insert(0, os.path.join(REF_ROOT, "_standin"))
    sys.path.insert(0, os.path.join(REF_ROOT, "ref"))
    import logging

    from environment.episode import Episodes   # the reference's, from baseline/_ref/ref
    from environment.tree import Tree
    from learn.rnad import RNaD
    from nn.net import MLP

    assert os.path.realpath(sys.modules["learn.rnad"].__file__).startswith(os.path.realpath(REF_ROOT))
    logging.disable(logging.INFO)
    cpu_dev = torch.device("cpu")
    tree = Tree(device=cpu_dev, max_actions=a, max_transitions=c, depth_bound=depth)
    for key, value in tables.items():
        setattr(tree, key, value)
    n_nodes = int(tree.index_tensor.shape[0])
    t_max = 2 * depth
    sample_batch = min(batch, args.reference_batch)
    torch.manual_seed(1234)
    net = MLP(a, 256, device=cpu_dev)

    # ---- env steps/s: Episodes.generate, the reference's own clock
    for _ in range(min(args.warmup, 2)):
        Episodes(tree, sample_batch).generate(net)
    total, elapsed = 0, 0.0
    for _ in range(args.steps):
        ep = Episodes(tree, sample_batch)
        ep.generate(net)
        elapsed += ep.generation_time
        total += int((ep.indices != 0).sum())
    value = total / elapsed

    # ---- updates/s: the rnad.py:495-526 loop body through the reference's own schedule loop
    n_updates = max(2, min(args.steps // 4, 5))
    trial = RNaD(tree=tree, device=cpu_dev, directory_name=f"bench_reference_{os.getpid()}", batch_size=sample_batch,
                 eta=0.2, lr=1e-3, gamma_averaging=0.01, logit_clip=2, b1_adam=0.0, bounds=[1], delta_m=[1],
                 net_params={"type": "MLP", "max_actions": a, "width": 256}, wandb=False)
    trial._RNaD__initialize()
    trial._RNaD__resume(checkpoint_mod=10 ** 9, expl_mod=10 ** 9, log_mod=10 ** 9)        # one warm-up update (m: 0 -> 1)
    trial.bounds, trial.delta_m = [2], [n_updates]
    t0 = time.perf_counter()
    trial._RNaD__resume(checkpoint_mod=10 ** 9, expl_mod=10 ** 9, log_mod=10 ** 9)
    learner_s = time.perf_counter() - t0
    import shutil

    shutil.rmtree(trial.directory, ignore_errors=True)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    sample = (f"{args.steps} x Episodes.generate of {sample_batch} games x {t_max} half-moves and {n_updates} learner "
              f"updates of {sample_batch} games, unmodified reference from baseline/_ref on {threads} torch threads")
    result = {
        "impl": "reference", "metric": "self_play_env_steps_per_sec", "value": value, "unit": "env_steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed * 1e3 / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args.config, depth, a, c, n_nodes, batch, t_max),
        "cpu_baseline": {"value": value, "unit": "env_steps/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "env_steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "learner": {"updates_per_sec": n_updates / learner_s, "ms_per_update": learner_s * 1e3 / n_updates,
                    "steps": n_updates, "batch": sample_batch,
                    "env_steps_per_sec": n_updates * sample_batch * t_max / learner_s,
                    "what": "reference RNaD.__resume loop body (rnad.py:495-526): Episodes.generate + Buffer.sample + "
                            "__learn + Adam + target-net average, device cpu"},
        "gpu_launches": 0,
    }
    print(json.dumps(result), flush=True)
    return result


def run_reference_port(args):
    """Fallback when baseline/_ref is absent (a checkout without /root/reference at build time): the CPU restatement
    of the rollout (oracle/) - `cpu_baseline.kind = "port"`."""
    from oracle import rnad_oracle as orc
    from nn.net import MLP

    depth, a, c, batch = CONFIGS[args.config]
    if args.batch:
        batch = args.batch
    if args.config in FAST_TREE_CONFIGS:
        tree = fast_tree(args.config, depth, a, c, torch.device("cpu"))   # on the host cores: minutes for the largest trees
    else:
        tree = make_tree(depth, a, c, seed=0)
    tables = {"index": tree.index_tensor, "value": tree.value_tensor, "chance": tree.chance_tensor,
              "expected_value": tree.expected_value_tensor, "legal": tree.legal_tensor}
    torch.manual_seed(1234)
    net = MLP(a, 256)
    weights = {k: v.detach().clone() for k, v in net.state_dict().items()}
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    t_max = 2 * depth
    sample_batch = min(batch, args.reference_batch)
    for _ in range(min(args.warmup, 2)):
        orc.rollout(tables, weights, sample_batch, t_max, seed=0)
    total, t0 = 0, time.perf_counter()
    for i in range(args.steps):
        out = orc.rollout(tables, weights, sample_batch, t_max, seed=i + 1)
        total += int((out["indices"] != 0).sum())
    elapsed = time.perf_counter() - t0
    value = total / elapsed
    world = int(os.environ.get("WORLD_SIZE", "1"))
    result = {
        "impl": "reference", "metric": "self_play_env_steps_per_sec", "value": value, "unit": "env_steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed * 1e3 / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args.config, depth, a, c, int(tree.index_tensor.shape[0]), batch, t_max),
        "cpu_baseline": {"value": value, "unit": "env_steps/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} rollouts of {sample_batch} games, CPU restatement of the reference "
                                   f"path (oracle/), torch-CPU ops on {threads} threads (baseline/_ref not installed)"},
        "e2e": {"value": value, "unit": "env_steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(result), flush=True)
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="games per GPU (default: the config's)")
    ap.add_argument("--precision", default="f16x2", choices=["tf32", "tf32x2", "f16x2", "fp32"])
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--reference-batch", type=int, default=65536)
    ap.add_argument("--learner-steps", type=int, default=200, help="timed learner updates (at least --steps)")
    ap.add_argument("--fp32-steps", type=int, default=10, help="timed rollouts of the fp32 engine (0 = skip)")
    ap.add_argument("--sustained-s", type=float, default=1.0, help="seconds of back-to-back rollouts (0 = skip)")
    ap.add_argument("--emit-tree", default="", help="(internal) write the configuration's tree tables to this file")
    args = ap.parse_args()
    if args.emit_tree:
        return emit_tree(args)
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
