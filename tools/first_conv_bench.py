"""Times the first conv layer (cin 1 / 2 -> 64 channels) at 4 x 1024^2 and 4 x 512^2; run once with PDA_FIRST_TC=0
(CUDA-core kernel) and once with PDA_FIRST_TC=1 (tensor-core kernel, the default)."""
import os, sys, torch
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


for (B, H) in ((4, 1024), (4, 512)):
    x0 = torch.randn(B, 1, H, H, generator=g).to(dev)
    x1 = torch.randn(B, 1, H, H, generator=g).to(dev)
    for cin in (1, 2):
        w = (torch.randn(64, cin, 3, 3, generator=g) * 0.3).to(dev)
        b = torch.randn(64, generator=g).to(dev)
        for dt in (torch.float16, torch.bfloat16):
            t = timed(lambda: ops.conv3x3_first(x0, x1 if cin == 2 else None, w, b, dtype=dt))
            gb = B * H * H * (128 + 4 * cin) / 1e9
            print(f"PDA_FIRST_TC={os.environ.get('PDA_FIRST_TC', '1')} {B}x{H}^2 cin={cin} {str(dt)[6:]}: {t*1e3:.1f} us  ({gb/t*1e3:.0f} GB/s)", flush=True)
