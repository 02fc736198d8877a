"""Achieved HBM bandwidth of the elementwise / first-layer kernels at the bench shapes (CUDA events, 20 launches)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilistic_domain_adaptation_b200 import ops
dev = torch.device("cuda:0")
def t(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("| kernel | shape | ms | algorithmic GB | GB/s | % of 6540 |"); print("|---|---|---|---|---|---|")
def row(name, shape, ms, gb): print(f"| {name} | {shape} | {ms:.3f} | {gb:.3f} | {gb/ms*1e3:.0f} | {100*gb/ms*1e3/6540.2:.0f} |")
for B, h, C in [(4, 128, 512), (4, 256, 256), (4, 512, 128), (4, 64, 512), (4, 128, 256), (4, 256, 128)]:
    x = torch.randn(B, h, h, C, device=dev).to(torch.bfloat16)
    ms = t(lambda: ops.upsample2x(x)); row("upsample2x", f"{B}x{h}x{h}x{C}", ms, x.numel() * 2 * 5 / 1e9)
    g = torch.randn(B, 2 * h, 2 * h, C, device=dev).to(torch.bfloat16)
    ms = t(lambda: ops.upsample2x_bwd(g)); row("upsample2x_bwd", f"{B}x{2*h}x{2*h}x{C}", ms, x.numel() * 2 * 5 / 1e9)
    del g
for B, H in [(4, 1024), (4, 512)]:
    img = torch.randn(B, 1, H, H, device=dev)
    w = torch.randn(64, 1, 3, 3, device=dev) * 0.1; b = torch.zeros(64, device=dev)
    ms = t(lambda: ops.conv3x3_first(img, None, w, b)); row("conv_first<1>", f"{B}x{H}x{H}", ms, B * H * H * (128 + 4) / 1e9)
    out = ops.conv3x3_first(img, None, w, b); dz = torch.randn_like(out)
    ms = t(lambda: ops.conv3x3_first_bwd(img, None, out, dz)); row("conv_first_bwd<1>", f"{B}x{H}x{H}", ms, B * H * H * (256 + 4) / 1e9)
    ms = t(lambda: ops.conv3x3_first_bwd(img, img, out, dz)); row("conv_first_bwd<2>", f"{B}x{H}x{H}", ms, B * H * H * (256 + 8) / 1e9)
    y = out; dfull = dz; dpool = torch.randn(B, H // 2, H // 2, 64, device=dev).to(torch.bfloat16)
    ms = t(lambda: ops.relu_pool_bwd(dfull, dpool, y)); row("relu_pool_bwd", f"{B}x{H}x{H}x64", ms, B * H * H * 64 * 2 * 3.25 / 1e9)
    del out, dz, dpool
