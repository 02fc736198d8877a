"""Fused Fcomb + consensus kernel alone: time vs number of samples S on 4 x 1024 x 1024 pixels of bf16 features.
Reports algorithmic HBM GB/s (140 B/px) and algorithmic TFLOP/s (2 * (4096 + S * 4160) FLOP/px)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilistic_domain_adaptation_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, H, W = 4, 1024, 1024
feat = torch.relu(torch.randn(B, H, W, 64, generator=g)).to(torch.bfloat16).to(dev)
w1 = (torch.randn(64, 70, 1, 1, generator=g) / 8).to(dev); b1 = torch.zeros(64, device=dev)
w2 = (torch.randn(64, 64, 1, 1, generator=g) / 8).to(dev); b2 = torch.zeros(64, device=dev)
w3 = (torch.randn(1, 64, 1, 1, generator=g)).to(dev); b3 = torch.zeros(1, device=dev)
print("| S | ms | algorithmic GB/s (140 B/px) | % of 6540 GB/s | algorithmic TFLOP/s | % of 1380 TFLOP/s |")
print("|---|---|---|---|---|---|")
SS = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32, 64]
for S in SS:
    z = torch.randn(S, B, 6, generator=g).to(dev)
    f = lambda: ops.fcomb_mc_consensus(feat, z, w1, b1, w2, b2, w3, b3, want_weight=False, want_mask=True)
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    n = 10
    for _ in range(n):
        f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    px = B * H * W
    gbs = px * 140 / (ms * 1e-3) / 1e9
    tf = px * 2.0 * (4096 + S * 4160) / (ms * 1e-3) / 1e12
    print(f"| {S} | {ms:.3f} | {gbs:.0f} | {100 * gbs / 6540.2:.1f} | {tf:.0f} | {100 * tf / 1379.8:.1f} |")
