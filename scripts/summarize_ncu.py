#!/usr/bin/env python
"""
Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the small, tracked
summaries under profiles/:  launch list -> per-kernel time shares;  full capture -> the
metrics DESIGN.md / bench.py quote plus the hottest SASS lines with their stall reasons.

    python scripts/summarize_ncu.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md
    python scripts/summarize_ncu.py kernel   gpurun_out/prof_rollout_r01.ncu-rep profiles/r01_rollout.md
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list: {src}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py` "
                f"(cold-cache, serialised replays: compare SHARES, not absolutes). {len(data)} launches, {total:.0f} us.\n\n")
        f.write("| us total | launches | share | kernel |\n|---:|---:|---:|---|\n")
        for name, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {t:.1f} | {c} | {100 * t / total:.1f}% | `{name[:110]}` |\n")
    print("wrote", dst)


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def kernel(rep, dst):
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full: {rep}\n\n")
        for r in raw[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"## `{name}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"| {m} | {r[i]} | {units[i]} |\n")
            f.write("\n")
        src = ncu_csv(rep, "source")
        hi = [i for i, r in enumerate(src) if "Source" in r]
        if hi:
            h = src[hi[0]]
            si, so = h.index("# Samples"), h.index("Source")
            stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
            data = [r for r in src[hi[0] + 1:] if len(r) > si and r[si].isdigit()]
            seen, uniq = set(), []
            for r in data:           # the source page repeats each line once per view
                key = (r[0], r[so])
                if key not in seen:
                    seen.add(key)
                    uniq.append(r)
            total = sum(int(r[si]) for r in uniq) or 1
            agg = collections.Counter()
            for r in uniq:
                for i in stall:
                    if r[i].isdigit():
                        agg[h[i]] += int(r[i])
            f.write(f"## warp-stall samples (first captured launch set, {total} samples)\n\n| reason | share |\n|---|---:|\n")
            for k, v in agg.most_common(8):
                f.write(f"| {k} | {100 * v / total:.1f}% |\n")
            f.write("\n## hottest SASS lines\n\n| samples | SASS | top stall |\n|---:|---|---|\n")
            for r in sorted(uniq, key=lambda r: -int(r[si]))[:20]:
                top = max(((int(r[i]) if r[i].isdigit() else 0, h[i]) for i in stall))
                f.write(f"| {r[si]} | `{r[so][:90]}` | {top[1]} ({top[0]}) |\n")
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
