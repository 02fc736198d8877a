"""Which Python call sites launch the small ATen kernels (fill / add / copy) of one eager mean-teacher training step?
Runs the step under torch.profiler with stacks and prints, per ATen op, the call sites inside this repo with counts."""
import collections
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from probabilistic_domain_adaptation_b200 import ProbabilisticUnet, consensus, steps  # noqa: E402
from probabilistic_domain_adaptation_b200.optim import FusedAdam  # noqa: E402
from probabilistic_domain_adaptation_b200.parallel import GradAllReducer  # noqa: E402

dev = torch.device("cuda:0")
model = ProbabilisticUnet(1, 1, [64, 128, 256, 512], 6, 3, 1.0, consensus_masking=True, rl_swap=True).to(dev).train()
teacher = copy.deepcopy(model)
for p in teacher.parameters():
    p.requires_grad = False
opt = FusedAdam(model.parameters(), lr=1e-5, capturable=True)
reducer = GradAllReducer(model)
ema = consensus.MomentumUpdater(model, teacher)
bp = steps.default_backprop(opt, reducer, model)
x1 = torch.randn(2, 1, 256, 256, device=dev)
x2 = x1 + 0.1
eps = torch.randn(16, 2, 6, device=dev)


def step():
    return steps.mean_teacher_step(model, teacher, opt, ema, x1, x2, n_samples=16, do_consensus_masking=True,
                                   backprop=bp, eps=eps)[0]


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA],
                            with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
sites = collections.Counter()
names = ("aten::fill_", "aten::zero_", "aten::add", "aten::add_", "aten::copy_", "aten::mul", "aten::neg", "aten::exp",
         "aten::zeros", "aten::zeros_like", "aten::_foreach_copy_", "aten::div", "aten::sum", "aten::mean")
for ev in prof.events():
    if ev.name in names and ev.device_time_total > 0 or ev.name in ("aten::fill_", "aten::add", "aten::add_"):
        site = "<autograd engine / no python frame>"
        for fr in ev.stack:
            if "probabilistic_domain_adaptation_b200" in fr or "tools/" in fr:
                site = fr.split("probabilistic_domain_adaptation_b200/")[-1]
                break
        sites[(ev.name, site)] += 1
for (name, site), n in sorted(sites.items(), key=lambda kv: -kv[1]):
    print(f"{n:4d}  {name:22s} {site}")
