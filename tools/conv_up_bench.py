"""Per-layer A/B of the fused up-sampling conv (ops.conv3x3_up) against upsample2x + conv3x3 at the up-path shapes of
the BASELINE inference workload (4 x 1024^2, no-grad activation format).  Prints ms per variant."""
import sys, torch
sys.path.insert(0, '.')
from probabilistic_domain_adaptation_b200 import ops

dev = torch.device('cuda:0')
dt = ops.INFER_DTYPE
g = torch.Generator().manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1024


def timed(fn, reps=10):
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


for (div, c0, c1, cout) in [(8, 512, 256, 256), (4, 256, 128, 128), (2, 128, 64, 64)]:
    h = S // div
    x_low = torch.randn(B, h, h, c0, generator=g).to(dev).to(dt)
    bridge = torch.randn(B, 2 * h, 2 * h, c1, generator=g).to(dev).to(dt)
    wt = (torch.randn(cout, c0 + c1, 3, 3, generator=g) * 0.02).to(dev)
    bias = torch.zeros(cout, device=dev)
    wp = ops.pack_conv3x3_weights(wt, dtype=dt)
    t_up = timed(lambda: ops.upsample2x(x_low))
    up = ops.upsample2x(x_low)
    t_conv = timed(lambda: ops.conv3x3(up, bridge, wp, bias))
    t_fused = timed(lambda: ops.conv3x3_up(x_low, bridge, wp, bias))
    flops = 2.0 * B * (2 * h) ** 2 * 9 * (c0 + c1) * cout
    print(f"{c0}+{c1}->{cout} @{2*h}: upsample {t_up:.3f} ms + conv {t_conv:.3f} ms ({flops/t_conv/1e9:.0f} TF/s) = "
          f"{t_up+t_conv:.3f} | fused {t_fused:.3f} ms ({flops/t_fused/1e9:.0f} TF/s)", flush=True)
