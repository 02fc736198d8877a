"""How long does the HOST need to enqueue one mean-teacher step / one MC inference step (no device sync inside)?
Compared with the device time of the same step this tells whether the step is launch-bound."""
import copy, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from probabilistic_domain_adaptation_b200 import consensus, steps
from probabilistic_domain_adaptation_b200.optim import FusedAdam
from probabilistic_domain_adaptation_b200.parallel import GradAllReducer

dev = torch.device("cuda:0")
model = bench.make_model(dev, consensus_masking=True, rl_swap=True).train()
teacher = copy.deepcopy(model)
for p in teacher.parameters():
    p.requires_grad = False
opt = FusedAdam(model.parameters(), lr=1e-5)
red = GradAllReducer(model)
ema = consensus.MomentumUpdater(model, teacher)
bp = steps.default_backprop(opt, red, model)
x1 = torch.randn(4, 1, 512, 512, device=dev)
x2 = x1 + 0.1
eps = torch.randn(16, 4, 6, device=dev)
def step():
    steps.mean_teacher_step(model, teacher, opt, ema, x1, x2, 16, True, backprop=bp, eps=eps)
for _ in range(3):
    step()
torch.cuda.synchronize()
for n in (1, 5):
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"train: {n} steps: host enqueue {1e3*(t1-t0)/n:.2f} ms/step, incl. device drain {1e3*(t2-t0)/n:.2f} ms/step")
m2 = bench.make_model(dev).eval()
xi = torch.randn(4, 1, 1024, 1024, device=dev)
for _ in range(3):
    consensus.sample_from_teacher(m2, xi, 16, do_consensus_masking=True, eps=eps)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    consensus.sample_from_teacher(m2, xi, 16, do_consensus_masking=True, eps=eps)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"infer: host enqueue {1e3*(t1-t0)/5:.2f} ms/step, incl. device drain {1e3*(t2-t0)/5:.2f} ms/step")
