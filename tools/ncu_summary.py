#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ (runs here, no GPU needed).
  python tools/ncu_summary.py launches gpurun_out/launches_infer.csv          -> per-kernel share table (markdown)
  python tools/ncu_summary.py full gpurun_out/prof_x.ncu-rep                  -> key metrics per captured launch
  python tools/ncu_summary.py stalls gpurun_out/prof_x.ncu-rep [N]            -> top-N source lines by stall samples
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__waves_per_multiprocessor",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__issue_active.avg.pct",
        "sm__inst_executed.sum", "smsp__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.avg.per_cycle_active", "sm__cycles_elapsed.max"]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hd = rows[h]
    ki, mv = hd.index("Kernel Name"), hd.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        try:
            t = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0]
        n, tot = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, tot + t)
    total = sum(t for _, t in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {name} | {n} | {t / 1e6:.3f} | {100 * t / total:.2f}% |")
    print(f"\ntotal {total / 1e6:.3f} ms over {sum(n for n, _ in agg.values())} launches")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hd, units = rows[0], rows[1]
    for r in rows[2:]:
        print("## launch")
        for name in ["Kernel Name", "Grid Size", "Block Size"]:
            print(f"- {name} = {r[hd.index(name)]}")
        for k in KEYS:
            for i, c in enumerate(hd):
                if c == k:
                    print(f"- {k} [{units[i]}] = {r[i]}")
        print()


def stalls(path, n=25):
    """Source lines of the FIRST captured kernel ranked by warp-stall samples, with the dominant stall reasons."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    hi = heads[0]
    end = heads[1] if len(heads) > 1 else len(rows)
    hd = rows[hi]
    si = hd.index("Warp Stall Sampling (All Samples)")
    reasons = [(i, c) for i, c in enumerate(hd) if c.startswith("stall_") and "Not Issued" not in c]
    lines, sass = [], []
    for r in rows[hi + 1:end]:
        if len(r) <= si or r[si] in ("-", ""):
            continue
        try:
            smp = float(r[si].replace(",", ""))
        except ValueError:
            continue
        rs = sorted(((float(r[i].replace(",", "") or 0), c[6:]) for i, c in reasons if r[i] not in ("-", "")),
                    reverse=True)[:3]
        rs = " ".join(f"{c}:{int(v)}" for v, c in rs if v > 0)
        if r[0]:
            lines.append((smp, r[0], r[1].strip()[:110], rs))
        else:
            sass.append((smp, r[2][-5:], r[3].strip()[:70], rs))
    tot = sum(x[0] for x in lines) or 1
    print(f"total samples {int(tot)}; top source lines:")
    for smp, ln, txt, rs in sorted(lines, key=lambda d: -d[0])[:n]:
        print(f"{100 * smp / tot:6.2f}%  L{ln:>4}  {txt}   [{rs}]")
    print("top SASS instructions:")
    for smp, ad, txt, rs in sorted(sass, key=lambda d: -d[0])[:n]:
        print(f"{100 * smp / tot:6.2f}%  {ad}  {txt}   [{rs}]")


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2])
    elif cmd == "full":
        full(sys.argv[2])
    else:
        stalls(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
