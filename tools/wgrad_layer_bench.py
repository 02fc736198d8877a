"""Per-layer timing of the weight-gradient kernel (memset + wgrad3x3_tc + scatter) at the shapes of the mean-teacher
training step (4 images of 512 x 512).  Prints a markdown table."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilistic_domain_adaptation_b200 import ops  # noqa: E402

LAYERS = [  # (name, H = W, c0, c1, cout)
    ("64->64 @512", 512, 64, 0, 64),
    ("64->128 @256", 256, 64, 0, 128),
    ("128->128 @256", 256, 128, 0, 128),
    ("128->256 @128", 128, 128, 0, 256),
    ("256->256 @128", 128, 256, 0, 256),
    ("256->512 @64", 64, 256, 0, 512),
    ("512->512 @64", 64, 512, 0, 512),
    ("512+256->256 @128", 128, 512, 256, 256),
    ("256+128->128 @256", 256, 256, 128, 128),
    ("128+64->64 @512", 512, 128, 64, 64),
]
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
print("| layer | GFLOP | us | TFLOP/s |\n|---|---|---|---|")
tot_t = tot_f = 0.0
for name, hw, c0, c1, cout in LAYERS:
    x = torch.randn(B, hw, hw, c0, generator=g).to(dev).to(torch.bfloat16)
    s1 = torch.randn(B, hw, hw, c1, generator=g).to(dev).to(torch.bfloat16) if c1 else None
    dz = torch.randn(B, hw, hw, cout, generator=g).to(dev).to(torch.bfloat16)
    f = lambda: ops.conv3x3_wgrad(x, s1, dz, want_bias=True)  # noqa: E731
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        f()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10 * 1e3
    flop = 2.0 * 9 * (c0 + c1) * cout * B * hw * hw
    tot_t += t
    tot_f += flop
    print(f"| {name} | {flop / 1e9:.0f} | {t:.1f} | {flop / t / 1e6:.0f} |", flush=True)
print(f"\nall ten: {tot_t:.0f} us, {tot_f / tot_t / 1e6:.0f} TFLOP/s")
