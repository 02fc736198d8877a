"""Where does the HOST time of a small-batch step go?  cProfile of the joint FixMatch step at LIVECell shape
(2 + 2 images of 256 x 256, launch-bound) and of MC inference on one 256 x 256 tile."""
import cProfile, io, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from probabilistic_domain_adaptation_b200 import consensus, steps
from probabilistic_domain_adaptation_b200.optim import FusedAdam
from probabilistic_domain_adaptation_b200.parallel import GradAllReducer

dev = torch.device("cuda:0")
model = bench.make_model(dev, consensus_masking=True, rl_swap=True).train()
opt = FusedAdam(model.parameters(), lr=1e-5)
red = GradAllReducer(model)
bp = steps.default_backprop(opt, red, model)
B, H = 2, 256
xs = torch.randn(B, 1, H, H, device=dev); ys = (torch.rand(B, 1, H, H, device=dev) > 0.5).float()
xt1 = torch.randn(B, 1, H, H, device=dev); xt2 = xt1 + 0.1
eps = torch.randn(16, B, 6, device=dev)
def step():
    steps.adamatch_step(model, opt, xs, ys, xt1, xt2, n_samples=16, do_consensus_masking=False, backprop=bp, eps=eps)
def infer():
    consensus.sample_from_teacher(model, xt1[:1], 16, do_consensus_masking=True, eps=eps[:, :1])
for name, fn, n in (("adamatch 2x256", step, 20), ("infer 1x256 S=16", infer, 50)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"== {name}: host enqueue {1e3*(t1-t0)/n:.3f} ms/step, with device drain {1e3*(t2-t0)/n:.3f} ms/step")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        fn()
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[:50]))
