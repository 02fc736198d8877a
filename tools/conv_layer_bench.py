"""Per-layer timing of the tensor-core conv3x3 at the shapes of one inference step (4 tiles of 1024 x 1024), for the
single-CTA kernel and the CTA-pair (cta_group::2) kernel.  Prints a markdown table.

    python tools/conv_layer_bench.py [--f16 0|1] [--pair 0,1]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilistic_domain_adaptation_b200 import _lib, ops  # noqa: E402

LAYERS = [  # (name, H = W, c0, c1, cout, pool)
    ("L0 64->64 @1024", 1024, 64, 0, 64, True),
    ("L1 64->128 @512", 512, 64, 0, 128, False),
    ("L1 128->128 @512", 512, 128, 0, 128, True),
    ("L2 128->256 @256", 256, 128, 0, 256, False),
    ("L2 256->256 @256", 256, 256, 0, 256, True),
    ("L3 256->512 @128", 128, 256, 0, 512, False),
    ("L3 512->512 @128", 128, 512, 0, 512, False),
    ("U2 512+256->256 @256", 256, 512, 256, 256, False),
    ("U1 256+128->128 @512", 512, 256, 128, 128, False),
    ("U0 128+64->64 @1024", 1024, 128, 64, 64, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--f16", type=int, default=1)
    ap.add_argument("--pair", default="0,1")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--size", type=int, default=1024, help="input tile size (the layer table is for 1024)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.load()
    dt = torch.float16 if args.f16 else torch.bfloat16
    modes = [int(m) for m in args.pair.split(",")]
    g = torch.Generator().manual_seed(0)
    print("| layer | GFLOP | " + " | ".join(f"pair={m}: us | TFLOP/s" for m in modes) + " |")
    print("|---|---|" + "---|---|" * len(modes))
    for name, hw, c0, c1, cout, pool in LAYERS:
        hw = hw * args.size // 1024
        name = name.split('@')[0] + f'@{hw}'
        s0 = torch.randn(args.batch, hw, hw, c0, generator=g).to(dev).to(dt)
        s1 = torch.randn(args.batch, hw, hw, c1, generator=g).to(dev).to(dt) if c1 else None
        w = (torch.randn(cout, c0 + c1, 3, 3, generator=g) * 0.02).to(dev)
        wp = ops.pack_conv3x3_weights(w, dtype=dt)
        bias = torch.zeros(cout, device=dev)
        flop = 2.0 * 9 * (c0 + c1) * cout * args.batch * hw * hw
        cells = []
        for m in modes:
            prev = lib.pda_set_conv_pair(m)
            f = lambda: ops.conv3x3(s0, s1, wp, bias, relu=True, want_full=True, want_pool=pool)  # noqa: E731
            for _ in range(3):
                f()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.reps):
                f()
            e1.record()
            torch.cuda.synchronize()
            lib.pda_set_conv_pair(prev)
            us = e0.elapsed_time(e1) / args.reps * 1e3
            cells.append(f"{us:.1f} | {flop / (us * 1e-6) / 1e12:.0f}")
        print(f"| {name} | {flop / 1e9:.0f} | " + " | ".join(cells) + " |")
        del s0, s1


if __name__ == "__main__":
    main()
