"""GPU: the same work through eager PyTorch (cuDNN / cuBLAS / ATen on the same B200) and through libpda_b200, timed side
by side.  SURVEY.md section 0 fact 10 / section 8(d): the honest on-box comparison for a reference that is pure PyTorch is
eager PyTorch itself.  The oracle port (oracle/punet_oracle.py, the reference's arithmetic op for op) plays the eager
arm, in fp32 (cuDNN TF32 convolutions, PyTorch's default) and under bf16 autocast (what torch_em's mixed-precision
trainer switches on); results are checked against each other and the timings are written to
gpurun_out/eager_compare.json.  Acts as a regression guard: the hand-written path must not be slower than eager."""
import copy
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def _record(key, value):
    path = os.path.join(ROOT, "gpurun_out", "eager_compare.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[key] = value
    json.dump(data, open(path, "w"), indent=1)


def _model_from(sd, dev, **kw):
    from probabilistic_domain_adaptation_b200.my_models import ProbabilisticUnet
    m = ProbabilisticUnet(1, 1, [64, 128, 256, 512], 6, 3, 1.0, **kw).to(dev)
    m.load_state_dict(sd)
    return m


def test_mc_inference_vs_eager_pytorch():
    from oracle import punet_oracle as po
    from probabilistic_domain_adaptation_b200 import consensus
    dev = torch.device("cuda:0")
    T, HW, S = 2, 1024, 16
    sd_cpu = po.make_state_dict(0, last_layer_gain=8.0)
    sd = {k: v.to(dev) for k, v in sd_cpu.items()}
    x, _, eps, _ = po.synthetic_inputs(T, HW, HW, s=S)
    x, eps = x.to(dev), eps.to(dev)
    model = _model_from(sd_cpu, dev).eval()

    def eager():
        with torch.no_grad():
            return po.sample_from_teacher(sd, x, eps, do_consensus_masking=True)

    def eager_bf16():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return po.sample_from_teacher(sd, x, eps, do_consensus_masking=True)

    def ours():
        with torch.no_grad():
            return consensus.sample_from_teacher(model, x, S, do_consensus_masking=True, eps=eps)

    y_ref = eager()[0]
    y_ours = ours()[0]
    # sanity only (last-layer gain 8, TF32 vs bf16 arithmetic); parity proper is tests/test_gpu_punet.py
    assert (y_ref - y_ours).abs().max() < 0.08 and (y_ref - y_ours).abs().mean() < 6e-3
    ms = {"eager_fp32_tf32": _time(eager, 3, 1), "eager_autocast_bf16": _time(eager_bf16, 3, 1), "ours": _time(ours, 10, 3)}
    work = T * HW * HW * S
    res = {k: {"ms_per_step": v, "px_samples_per_s": work / (v * 1e-3)} for k, v in ms.items()}
    res["workload"] = f"{T} tiles 1x{HW}x{HW}, S={S}, forward + S samples + consensus mask, device-resident"
    res["speedup_vs_eager_fp32"] = ms["eager_fp32_tf32"] / ms["ours"]
    res["speedup_vs_eager_autocast_bf16"] = ms["eager_autocast_bf16"] / ms["ours"]
    _record("mc_inference", res)
    print("\nMC inference", json.dumps(res))
    assert ms["ours"] < min(ms["eager_fp32_tf32"], ms["eager_autocast_bf16"])


def test_mean_teacher_step_vs_eager_pytorch():
    from oracle import punet_oracle as po
    from probabilistic_domain_adaptation_b200 import consensus, steps
    from probabilistic_domain_adaptation_b200.optim import FusedAdam
    dev = torch.device("cuda:0")
    B, HW, S = 4, 512, 16
    sd_cpu = po.make_state_dict(0, last_layer_gain=8.0)
    x, _, eps, eps_post = po.synthetic_inputs(B, HW, HW, s=S)
    x1, x2, eps, eps_post = (x + 0.1).to(dev), (x - 0.1).to(dev), eps.to(dev), eps_post.to(dev)

    student = {k: v.to(dev).clone().requires_grad_(True) for k, v in sd_cpu.items()}
    teacher = {k: v.detach().clone() for k, v in student.items()}
    opt = torch.optim.Adam(list(student.values()), lr=1e-5, fused=True)

    def eager_step(autocast):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=False):
            y, z, _ = po.sample_from_teacher(teacher, x1, eps, do_consensus_masking=True)   # fp32, as the trainers do
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = po.training_loss(student, x2, y, eps_post, z, beta=1.0, consensus_masking=True, rl_swap=True)
        out["loss"].backward()
        opt.step()
        with torch.no_grad():
            teacher.update(po.momentum_update(teacher, {k: v.detach() for k, v in student.items()}, 0.999))

    model = _model_from(sd_cpu, dev, consensus_masking=True, rl_swap=True).train()
    tmodel = copy.deepcopy(model)
    for p in tmodel.parameters():
        p.requires_grad = False
    fopt = FusedAdam(model.parameters(), lr=1e-5)
    ema = consensus.MomentumUpdater(model, tmodel)
    bp = steps.default_backprop(fopt, None, model)

    def ours():
        steps.mean_teacher_step(model, tmodel, fopt, ema, x1, x2, S, True, backprop=bp, eps=eps)

    ms = {"eager_fp32_tf32": _time(lambda: eager_step(False), 3, 1),
          "eager_autocast_bf16_student": _time(lambda: eager_step(True), 3, 1), "ours": _time(ours, 10, 3)}
    res = {k: {"ms_per_step": v, "img_per_s": B / (v * 1e-3)} for k, v in ms.items()}
    res["workload"] = f"mean-teacher consensus-masking step, {B}x1x{HW}x{HW}, S={S} (teacher MC + student fwd/bwd + Adam + EMA)"
    res["speedup_vs_eager_fp32"] = ms["eager_fp32_tf32"] / ms["ours"]
    res["speedup_vs_eager_autocast_bf16"] = ms["eager_autocast_bf16_student"] / ms["ours"]
    _record("mean_teacher_step", res)
    print("\nmean-teacher step", json.dumps(res))
    assert ms["ours"] < min(ms["eager_fp32_tf32"], ms["eager_autocast_bf16_student"])
