"""gpurun_out/traffic.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum of ONE
inference step, tools/gpu_traffic.sh) -> profiles/dram_traffic.json (read by bench.py for roofline.traffic) and a
markdown table per launch.

    python tools/ncu_traffic_summary.py gpurun_out/traffic.csv <build tag> [tiles tile samples]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAYER_SHAPES = None


def main():
    path, tag = sys.argv[1], sys.argv[2]
    T, HW, S = (int(a) for a in sys.argv[3:6]) if len(sys.argv) >= 6 else (4, 1024, 16)
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hd = rows[h]
    ki, mn, mv, mu, idc = (hd.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    launches = {}
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        d = launches.setdefault(int(r[idc]), {"kernel": r[ki].split("(")[0].replace("void ", "").replace("pda::", "")})
        v = float(r[mv].replace(",", ""))
        unit = r[mu]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
        d[r[mn]] = v * scale
    conv = [d for d in launches.values() if d["kernel"].startswith("conv3x3_tc")]
    fc = [d for d in launches.values() if d["kernel"].startswith("fcomb_tc")]

    def traffic(d):
        return d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)

    # algorithmic bytes of the conv launches of one inference step: activations in + out (16-bit) + weights, per layer
    px = T * HW * HW
    nets = []
    for _ in range(2):  # prior encoder, U-Net contracting path share the shapes
        nets += [(64, 64, 0), (64, 64, 0), (64, 128, 1), (128, 128, 1), (128, 128, 1), (128, 256, 2), (256, 256, 2),
                 (256, 256, 2), (256, 512, 3), (512, 512, 3), (512, 512, 3)]
    nets += [(256, 256, 2), (256, 256, 2), (128, 128, 1), (128, 128, 1), (64, 64, 0), (64, 64, 0)]
    alg = sum((cin + cout) * 2.0 * px / 4 ** lvl + 9.0 * cin * cout * 2 for cin, cout, lvl in nets)
    # first conv of an up block with the bilinear x2 fused in: the low-resolution tensor (c_low channels at level
    # lvl + 1) and the bridge (c_br channels at level lvl) are read, cout channels written
    fused = [(512, 256, 256, 2), (256, 128, 128, 1), (128, 64, 64, 0)]
    alg += sum(c_low * 2.0 * px / 4 ** (lvl + 1) + (c_br + cout) * 2.0 * px / 4 ** lvl + 9.0 * (c_low + c_br) * cout * 2
               for c_low, c_br, cout, lvl in fused)
    nets = nets + fused
    out = {
        "build": tag, "workload": {"tiles": T, "tile": HW, "samples": S},
        "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on one inference step "
               "(tools/gpu_traffic.sh); per launch = mean over the launches of the kernel family",
        "conv3x3_tc": {"launches": len(conv), "dram_bytes_per_launch": sum(traffic(d) for d in conv) / max(1, len(conv)),
                       "algorithmic_bytes_per_launch": alg / max(1, len(nets)), "dram_bytes_per_step": sum(traffic(d) for d in conv),
                       "algorithmic_bytes_per_step": alg},
        "fcomb_tc": {"launches": len(fc), "dram_bytes_per_launch": sum(traffic(d) for d in fc) / max(1, len(fc)),
                     "algorithmic_bytes_per_launch": 140.0 * px},
    }
    with open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(f"# {tag}: DRAM traffic per launch of one inference step ({T} tiles {HW}x{HW}, S={S})\n")
    print("| # | kernel | read MB | write MB | us |\n|---|---|---|---|---|")
    for i, (k, d) in enumerate(sorted(launches.items())):
        print(f"| {i} | {d['kernel']} | {d.get('dram__bytes_read.sum', 0) / 1e6:.1f} | "
              f"{d.get('dram__bytes_write.sum', 0) / 1e6:.1f} | {d.get('gpu__time_duration.sum', 0):.1f} |")
    print(f"\nconv launches: {len(conv)}; DRAM {out['conv3x3_tc']['dram_bytes_per_step'] / 1e9:.2f} GB per step against "
          f"{alg / 1e9:.2f} GB algorithmic (activations in + out + weights); fcomb "
          f"{out['fcomb_tc']['dram_bytes_per_launch'] / 1e6:.0f} MB against {140.0 * px / 1e6:.0f} MB.")


if __name__ == "__main__":
    main()
