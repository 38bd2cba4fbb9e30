#!/usr/bin/env python
"""
Standalone K1 numbers (north_star: "achieved HBM GB/s for the gather path"): `observe_kernel` and `step_kernel`
(csrc/env.cu; States.observations / States.step, reference episode.py:46-68, 84-125) through the C ABI on the cfg3 tree
(max_actions 3, max_transitions 3, depth 6: 14.9 M nodes, 7 GB of packed tables - HBM-resident, no reuse between
games), node ids taken from a real rollout at every depth.  CUDA events around each launch, a 256 MiB write between
launches (cold L2).  Prints one JSON object; run under ncu for the counters (scripts/gpu_r02*.sh).

Algorithmic bytes per game and call:
  observe  idx 4 + node record 4A^2 + 4 read, observation 8A^2 + mask 4A written
  step     idx 4 + two actions 16 + transition entry 12C read, idx 4 + reward 4 written
"""
import argparse
import ctypes
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "r-nad_b200"), REPO):
    sys.path.insert(0, p)

import torch  # noqa: E402

import _b200  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0, help="games (default: the configuration's)")
    args = ap.parse_args()
    from environment.episode import Episodes
    from nn.net import MLP

    depth, a, c, batch = bench.CONFIGS[args.config]
    batch = args.batch or batch
    dev = torch.device("cuda")
    tree = bench.fast_tree(args.config, depth, a, c, dev) if args.config in bench.FAST_TREE_CONFIGS else None
    if tree is None:
        tree = bench.make_tree(depth, a, c)
        tree.to(dev)
    packed = tree.packed()
    torch.manual_seed(0)
    net = MLP(a, 256, device=dev)
    ep = Episodes(tree, batch)
    ep.generate(net)
    L = _b200.lib()
    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    obs = torch.empty((batch, 2, a, a), dtype=torch.float32, device=dev)
    mask = torch.empty((batch, a), dtype=torch.float32, device=dev)
    rewards = torch.empty(batch, dtype=torch.float32, device=dev)
    alive = torch.zeros(1, dtype=torch.int32, device=dev)
    bytes_observe = 4 + 4 * a * a + 4 + 8 * a * a + 4 * a
    bytes_step = 4 + 16 + 12 * c + 4 + 4

    def timed(fn):
        best = []
        for _ in range(args.reps):
            flush_buf.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            best.append(e0.elapsed_time(e1))
        return sum(best) / len(best), min(best)

    rows = []
    for t in range(0, ep.t_eff + 1, 2):          # full moves: row half-move t, column half-move t + 1
        idx = ep.indices[t].to(torch.int32).contiguous()
        row_a = ep.actions[t].argmax(-1).contiguous()
        col_a = ep.actions[t + 1].argmax(-1).contiguous()
        n_alive = int((idx != 0).sum())
        mean_o, min_o = timed(lambda: L.rnad_observe(_b200.ptr(packed.ev_tab), a, packed.S, _b200.ptr(idx), 0, batch,
                                                     _b200.ptr(obs), _b200.ptr(mask), _b200.stream()))
        work = idx.clone()

        def step():
            work.copy_(idx)
            L.rnad_step(_b200.ptr(packed.tr_tab), a, c, _b200.ptr(work), _b200.ptr(row_a), _b200.ptr(col_a), None,
                        ctypes.c_uint64(7), t + 1, 0, batch, _b200.ptr(rewards), _b200.ptr(alive), _b200.stream())

        def copy_only():
            work.copy_(idx)

        mean_s, _ = timed(step)
        mean_c, _ = timed(copy_only)
        mean_s = max(mean_s - mean_c, 1e-4)
        rows.append({"half_move": t, "distinct_nodes": int(idx.unique().numel()), "alive": n_alive,
                     "observe_us": round(mean_o * 1e3, 2), "observe_GBps": round(bytes_observe * batch / mean_o / 1e6, 1),
                     "step_us": round(mean_s * 1e3, 2), "step_GBps": round(bytes_step * batch / mean_s / 1e6, 1)})
    peaks = bench.measured_peaks()
    out = {"config": args.config, "nodes": packed.S, "packed_table_bytes": packed.nbytes(), "batch": batch,
           "algorithmic_bytes_per_game": {"observe": bytes_observe, "step": bytes_step},
           "hbm_peak_GBps": peaks["hbm_gbs"], "levels": rows}
    deepest = rows[-1]
    out["deepest_level"] = {"observe_frac_of_hbm": round(deepest["observe_GBps"] / peaks["hbm_gbs"], 3),
                            "step_frac_of_hbm": round(deepest["step_GBps"] / peaks["hbm_gbs"], 3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
