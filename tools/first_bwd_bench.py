"""Times the first-layer weight-gradient kernels (cin 1 and 2) and the bilinear up-sampling backward at the training
bench's shapes (4 x 512^2)."""
import sys, torch
sys.path.insert(0, '.')
from probabilistic_domain_adaptation_b200 import ops
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


B, H = 4, 512
x0 = torch.randn(B, 1, H, H, generator=g).to(dev)
x1 = torch.randn(B, 1, H, H, generator=g).to(dev)
dout = torch.randn(B, H, H, 64, generator=g).to(dev).to(torch.bfloat16)
out = torch.randn(B, H, H, 64, generator=g).to(dev).to(torch.bfloat16)
for cin, xb in ((1, None), (2, x1)):
    for pre in (True, False):
        t = timed(lambda: ops.conv3x3_first_bwd(x0, xb, None if pre else out, dout, premasked=pre))
        print(f"conv_first_bwd cin={cin} premasked={pre}: {t*1e3:.1f} us", flush=True)
for (h, c) in ((256, 128), (128, 256), (64, 512)):
    d = torch.randn(B, 2 * h, 2 * h, c, generator=g).to(dev).to(torch.bfloat16)
    t = timed(lambda: ops.upsample2x_bwd(d))
    gb = (d.numel() * 2 * 1.25) / 1e9
    print(f"upsample2x_bwd {c}ch -> {h}^2: {t*1e3:.1f} us ({gb/t*1e3:.0f} GB/s algorithmic)", flush=True)
