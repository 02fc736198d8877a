#!/usr/bin/env python
"""Benchmark of the PUNet Monte-Carlo inference hot path (BASELINE.json metric: px*samples/s at S=16).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--tiles T] [--size HW] [--samples S]

One "step" = one batch of T synthetic HxW tiles through  forward(prior + U-Net) + S-sample fused
Fcomb/sigmoid/mean/consensus-mask  (the path of mean_teacher_trainer.py:72-88 / punet_predictions.py:29-33).
Under torchrun every rank processes its own tiles (weak scaling, no data-path collective); the time is the
max over ranks of the CUDA-event time of the K timed steps.

`--impl reference` times the reference's CPU implementation of the same path (the oracle port of the
reference PyTorch code: the reference tree itself does not exist on the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PX_FORWARD = 2_509_056  # prior 701 568 + U-Net 1 807 488, incl. the cin=1 first layers
FIRST_LAYER_FLOP_PER_PX = 2 * 2 * 9 * 64  # the two cin=1 first layers run on CUDA cores (HBM-bound)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tiles", type=int, default=4, help="tiles per GPU per step")
    ap.add_argument("--size", type=int, default=1024, help="tile height = width")
    ap.add_argument("--samples", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="both", choices=["infer", "train", "both"],
                    help="infer: MC inference leg only; train: mean-teacher step leg only")
    ap.add_argument("--train-batch", type=int, default=4, help="images per GPU per mean-teacher step (MitoEM: 4)")
    ap.add_argument("--train-size", type=int, default=512)
    ap.add_argument("--bucket-mb", type=float, default=25.0, help="gradient all-reduce bucket size (training legs)")
    ap.add_argument("--no-graph", action="store_true", help="training legs: time the Python-launched step only")
    ap.add_argument("--reserve-sms", type=int, default=0,
                    help="N > 1: SMs the persistent kernels leave to NCCL so that the all-reduce overlaps the backward")
    ap.add_argument("--grad-comm", default="fp32", choices=["fp32", "bf16"], help="wire format of the gradient all-reduce")
    ap.add_argument("--no-extras", action="store_true", help="skip the S sweep and the source-training step")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return p.get("hbm_gbs", 6650.0), p.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        clocks, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                clocks.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        clocks.sort()
        med = clocks[len(clocks) // 2] if clocks else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(clocks)}


def cpu_reference_run(size, samples, tiles, steps, warmup):
    """The reference path on host cores (oracle port), fp32, all threads.  Returns (px*samples/s, s/step, cores)."""
    import torch
    from oracle import punet_oracle as po
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = po.make_state_dict(0, last_layer_gain=8.0)
    x, _, eps, _ = po.synthetic_inputs(tiles, size, size, s=samples)

    def step():
        with torch.no_grad():
            y, z, _ = po.sample_from_teacher(sd, x, eps, do_consensus_masking=True)
        return y, z

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return tiles * size * size * samples / dt, dt, cores


def cpu_reference_train_step(batch, size, samples):
    """One mean-teacher consensus step of the reference on host cores (oracle port, fp32 autograd + torch Adam):
    teacher MC + consensus mask, student Dice-ELBO + L2, backward, Adam, EMA.  Returns (img/s, s/step, cores)."""
    import torch
    from oracle import punet_oracle as po
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    student = {k: v.clone().requires_grad_(True) for k, v in po.make_state_dict(0, last_layer_gain=8.0).items()}
    teacher = {k: v.detach().clone() for k, v in student.items()}
    opt = torch.optim.Adam(list(student.values()), lr=1e-5)
    x, _, eps, eps_post = po.synthetic_inputs(batch, size, size, s=samples)
    x1, x2 = x + 0.1, x - 0.1

    def step():
        with torch.no_grad():
            y, z, _ = po.sample_from_teacher(teacher, x1, eps, do_consensus_masking=True)
        opt.zero_grad()
        out = po.training_loss(student, x2, y, eps_post, z, beta=1.0, consensus_masking=True, rl_swap=True)
        out["loss"].backward()
        opt.step()
        with torch.no_grad():
            new = po.momentum_update(teacher, {k: v.detach() for k, v in student.items()}, 0.999)
            teacher.update(new)

    step()
    t0 = time.perf_counter()
    step()
    dt = time.perf_counter() - t0
    return batch / dt, dt, cores


def eager_gpu_baseline(dev, T, HW, S, steps=3, warmup=1):
    """The same inference workload through EAGER PyTorch on the same B200 (cuDNN / cuBLAS / ATen kernels): the oracle
    port is the reference's arithmetic op for op (the reference's GPU path IS eager PyTorch).
    fp32 with cuDNN's default TF32 convolutions, and under bf16 autocast.  A reported baseline (GPU vs GPU), measured
    inside the default run so that the driver's record carries it; never part of the product path."""
    import torch
    from oracle import punet_oracle as po
    sd = {k: v.to(dev) for k, v in po.make_state_dict(0, last_layer_gain=8.0).items()}
    x, _, eps, _ = po.synthetic_inputs(T, HW, HW, s=S)
    x, eps = x.to(dev), eps.to(dev)

    def run(autocast):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            return po.sample_from_teacher(sd, x, eps, do_consensus_masking=True)

    out = {}
    for name, ac in (("eager_fp32_tf32", False), ("eager_autocast_bf16", True)):
        for _ in range(warmup):
            run(ac)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            run(ac)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"ms_per_step": ms, "value": float(T) * HW * HW * S / (ms * 1e-3), "unit": "px*samples/s"}
    out["sample"] = f"{T} tiles 1x{HW}x{HW}, S={S}, {steps} timed steps after {warmup} warm-up, device-resident"
    out["what"] = "oracle port of the reference run as eager PyTorch on cuda:0 (cuDNN conv, ATen elementwise)"
    del sd, x, eps
    torch.cuda.empty_cache()
    return out


def load_traffic_profile(T, HW, S):
    """DRAM bytes per launch of the dominant kernels from the ncu capture of THIS workload, written by
    tools/ncu_traffic_summary.py into profiles/dram_traffic.json (which names the build it was taken on)."""
    path = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        prof = json.load(fh)
    if prof.get("workload") != {"tiles": T, "tile": HW, "samples": S}:
        return None
    return prof


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: one 512x512 tile per step keeps `--steps 10 --warmup 3` within a few minutes
    size, tiles = min(args.size, 512), 1
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    val, dt, cores = cpu_reference_run(size, args.samples, tiles, steps, warmup)
    sample = f"{tiles} tile {size}x{size}, S={args.samples}, {steps} timed steps after {warmup} warm-up"
    print(json.dumps({
        "impl": "reference", "metric": "punet_mc_px_samples_per_s", "value": val, "unit": "px*samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"PUNet MC inference S={args.samples} + consensus mask, {args.size}x{args.size} tiles",
                   "timed_sample": sample},
        "cpu_baseline": {"value": val, "unit": "px*samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "px*samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))



def make_model(dev, **kw):
    """Random-init PUNet of the scripts' architecture (reference init scheme, seed 0); the last Fcomb layer is scaled
    so that the sampled probabilities spread over (0, 1) and the consensus mask takes both values."""
    import torch
    from probabilistic_domain_adaptation_b200 import ProbabilisticUnet
    torch.manual_seed(0)
    m = ProbabilisticUnet(1, 1, [64, 128, 256, 512], 6, 3, 1.0, **kw).to(dev)
    with torch.no_grad():
        m.fcomb.last_layer.weight.mul_(8.0)
    return m


TRAIN_FLOP_PER_PX = 3 * 3_229_056 + 2_509_056  # student fwd+bwd (3x forward) + teacher forward


def run_train_leg(args, dev, world, rank, local, lib, barrier):
    """Mean-teacher consensus step (BASELINE config 3, mean_teacher_trainer.py:101-131): teacher forward + S prior
    samples + consensus mask, student forward(x2, y) + Dice-ELBO(y, z) + L2 + backward, gradient all-reduce,
    Adam, EMA.  Returns the "train" object of the JSON line (rank 0) or None."""
    import copy
    import torch
    import torch.distributed as dist
    from probabilistic_domain_adaptation_b200 import ProbabilisticUnet, consensus, ops, steps
    from probabilistic_domain_adaptation_b200.optim import FusedAdam
    from probabilistic_domain_adaptation_b200.parallel import GradAllReducer

    Bt, HW, S = args.train_batch, args.train_size, args.samples
    model = make_model(dev, consensus_masking=True, rl_swap=True).train()
    teacher = copy.deepcopy(model)
    for p in teacher.parameters():
        p.requires_grad = False
    opt = FusedAdam(model.parameters(), lr=1e-5, capturable=True)
    import torch as _t
    comm = _t.bfloat16 if args.grad_comm == "bf16" else None
    reducer = GradAllReducer(model, bucket_mb=args.bucket_mb, reserve_sms=args.reserve_sms, comm_dtype=comm)
    ema = consensus.MomentumUpdater(model, teacher)
    backprop = steps.default_backprop(opt, reducer, model)
    use_graph = not args.no_graph
    g = torch.Generator().manual_seed(11 + rank)
    x = torch.randn(Bt, 1, HW, HW, generator=g)
    host_x1 = (x + 0.1 * torch.randn(Bt, 1, HW, HW, generator=g)).pin_memory()
    host_x2 = (x + 0.25 * torch.randn(Bt, 1, HW, HW, generator=g)).pin_memory()
    x1, x2 = host_x1.to(dev), host_x2.to(dev)
    eps = torch.randn(S, Bt, 6, generator=torch.Generator().manual_seed(3)).to(dev)
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()

    def step_fn(a, b):
        return steps.mean_teacher_step(model, teacher, opt, ema, a, b, n_samples=S, do_consensus_masking=True,
                                       backprop=backprop, eps=eps)[0]

    def step_resident():
        return step_fn(x1, x2)

    def step_e2e():
        a = host_x1.to(dev, non_blocking=True)
        b = host_x2.to(dev, non_blocking=True)
        loss = steps.mean_teacher_step(model, teacher, opt, ema, a, b, n_samples=S, do_consensus_masking=True,
                                       backprop=backprop, eps=eps)[0]
        host_loss.copy_(loss.detach(), non_blocking=True)
        return loss

    def timed(fn, steps_, profile=False):
        barrier()
        ops.PROFILE = [] if profile else None
        lib.pda_reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps_):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.pda_launch_count()
        prof, ops.PROFILE = ops.PROFILE, None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, launches, prof

    for _ in range(args.warmup):
        loss = step_resident()
    # pass 1, launched kernel by kernel from Python: per-kernel CUDA-event profile and launch count
    ms_launch, launches, prof = timed(step_resident, args.steps, profile=True)
    ms, ms_graph = ms_launch, None
    if use_graph:
        # pass 2: the same step body captured once into a CUDA graph and replayed (steps.GraphedStep): the public
        # API for fixed-shape training; removes the host from the step
        gstep = steps.GraphedStep(lambda a, b: step_fn(a, b), (x1, x2), optimizer=opt, warmup=1)

        def step_graph():
            return gstep(x1, x2)

        def step_e2e_graph():
            loss = gstep(host_x1, host_x2)   # H2D copies of both views into the static inputs, then the replay
            host_loss.copy_(loss.detach(), non_blocking=True)
            return loss
        for _ in range(args.warmup):
            loss = step_graph()
        ms_graph, _, _ = timed(step_graph, args.steps)
        ms = ms_graph
        for _ in range(max(1, args.warmup // 2)):
            step_e2e_graph()
        ms_e2e, _, _ = timed(step_e2e_graph, args.steps)
    else:
        for _ in range(max(1, args.warmup // 2)):
            step_e2e()
        ms_e2e, _, _ = timed(step_e2e, args.steps)
    final_loss = float(loss.item())
    reducer.remove()

    # BASELINE config 2: source training step (punet_trainer.py:24-36) on LIVECell shape 8x1x512x512
    src = None
    if not args.no_extras:
        Bs = 8
        xs = torch.randn(Bs, 1, HW, HW, generator=g).to(dev)
        ys = (torch.rand(Bs, 1, HW, HW, generator=g) > 0.5).float().to(dev)
        red2 = GradAllReducer(model, reserve_sms=args.reserve_sms, comm_dtype=comm)
        bp2 = steps.default_backprop(opt, red2, model)
        for _ in range(2):
            steps.punet_step(model, opt, xs, ys, backprop=bp2)
        ms_src, _, _ = timed(lambda: steps.punet_step(model, opt, xs, ys, backprop=bp2), max(3, args.steps // 2))
        red2.remove()
        nsrc = max(3, args.steps // 2)
        src = {"metric": "punet_source_train_img_per_s", "value": Bs * world * nsrc / (ms_src * 1e-3), "unit": "img/s",
               "ms_per_step": ms_src / nsrc,
               "config": {"workload": f"source ELBO step (Dice + KL + L2, Adam) on {Bs}x1x{HW}x{HW} per GPU "
                                      f"(BASELINE config 2, LIVECell shape)", "parallelism": f"dp{world}"}}
    # BASELINE config 4: joint FixMatch ("AdaMatch", adamatch_trainer.py:62-102) with consensus weighting, data parallel:
    # source ELBO + weak-view MC pseudo-labels (the model itself) + strong-view target ELBO, one backward, all-reduce, Adam
    joint = None
    if not args.no_extras:
        joint = {}
        for tag, (Bj, Hj) in {"livecell_2x256": (2, 256), "mitoem_4x512": (4, 512)}.items():
            xs = torch.randn(Bj, 1, Hj, Hj, generator=g).to(dev)
            ys = (torch.rand(Bj, 1, Hj, Hj, generator=g) > 0.5).float().to(dev)
            xt = torch.randn(Bj, 1, Hj, Hj, generator=g)
            xt1 = (xt + 0.1 * torch.randn(Bj, 1, Hj, Hj, generator=g)).to(dev)
            xt2 = (xt + 0.25 * torch.randn(Bj, 1, Hj, Hj, generator=g)).to(dev)
            epsj = torch.randn(S, Bj, 6, generator=torch.Generator().manual_seed(3)).to(dev)
            # consensus WEIGHTING (--consensus without --masking, livecell_adamatch.py:122,150): the model keeps
            # consensus_masking=True (= "use consm"), the trainer returns fp32 k/16 weights instead of the int64 mask
            red3 = GradAllReducer(model, reserve_sms=args.reserve_sms, comm_dtype=comm)
            bp3 = steps.default_backprop(opt, red3, model)
            fnj = lambda: steps.adamatch_step(model, opt, xs, ys, xt1, xt2, n_samples=S, do_consensus_masking=False,  # noqa: E731
                                              backprop=bp3, eps=epsj)
            for _ in range(2):
                fnj()
            nj = max(3, args.steps // 2)
            ms_j, launches_j, _ = timed(fnj, nj)
            ms_j_launch = ms_j
            if use_graph:
                gj = steps.GraphedStep(lambda a, b, c, d: steps.adamatch_step(
                    model, opt, a, b, c, d, n_samples=S, do_consensus_masking=False, backprop=bp3, eps=epsj)[0],
                    (xs, ys, xt1, xt2), optimizer=opt, warmup=1)
                gj(xs, ys, xt1, xt2)
                ms_j, _, _ = timed(lambda: gj(xs, ys, xt1, xt2), nj)
                del gj
            red3.remove()
            joint[tag] = {"metric": "adamatch_joint_train_img_per_s", "value": 2 * Bj * world * nj / (ms_j * 1e-3),
                          "unit": "img/s (source + target)", "ms_per_step": ms_j / nj,
                          "ms_per_step_python_launched": ms_j_launch / nj, "cuda_graph": use_graph,
                          "launches_per_step": launches_j / nj,
                          "config": {"workload": f"joint FixMatch step with consensus weighting: {Bj} source + {Bj} target "
                                                 f"images of 1x{Hj}x{Hj} per GPU, S={S} (BASELINE config 4)",
                                     "parallelism": f"dp{world}"}}
    if rank != 0:
        return None
    cpu_train = None
    if world == 1 and not args.no_cpu_baseline:
        cval, cdt, cores = cpu_reference_train_step(1, 256, S)
        cpu_train = {"value": cval, "unit": "img/s", "cores": cores, "kind": "port",
                     "sample": f"1 image 256x256 (1/16 of the pixels of a {HW}x{HW} image), S={S}, 1 timed step "
                               f"after 1 warm-up ({cdt:.1f} s); per-pixel cost is size-independent",
                     "value_at_bench_shape": cval * (256.0 * 256.0) / (HW * HW)}
    _, tf_peak, peak_src = load_peaks()
    kinds = {}
    for k, a, b, w in prof:
        t, f, n = kinds.get(k, (0.0, 0.0, 0))
        kinds[k] = (t + a.elapsed_time(b), f + w, n + 1)
    tc_ms = sum(kinds.get(k, (0, 0, 0))[0] for k in ("conv3x3_tc", "wgrad3x3_tc"))
    tc_flop = sum(kinds.get(k, (0, 0, 0))[1] for k in ("conv3x3_tc", "wgrad3x3_tc"))
    achieved = tc_flop / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    per_kernel = {k: {"ms_per_step": t / args.steps, "launches_per_step": n / args.steps,
                      "tflops": (f / (t * 1e-3) / 1e12 if k in ("conv3x3_tc", "wgrad3x3_tc") and t > 0 else None)}
                  for k, (t, f, n) in kinds.items()}
    imgs = float(Bt) * world * args.steps
    return {
        "metric": "mt_consensus_train_img_per_s", "value": imgs / (ms * 1e-3), "unit": "img/s",
        "ms_per_step": ms / args.steps,
        "config": {"workload": f"mean-teacher consensus-masking step: teacher forward + S={S} samples + mask, student "
                               f"Dice-ELBO fwd/bwd + L2, grad all-reduce, Adam, EMA on {Bt}x1x{HW}x{HW} per GPU "
                               f"(BASELINE config 3, MitoEM shape)",
                   "batch_per_gpu": Bt, "patch": HW, "samples": S, "parallelism": f"dp{world}",
                   "grad_allreduce": (f"NCCL, {args.grad_comm} on the wire, {args.bucket_mb:g} MB buckets, "
                                      f"{args.reserve_sms} SMs left to NCCL") if world > 1 else None},
        "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "img/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": 2 * host_x1.numel() * 4, "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches), "final_loss": final_loss,
        "cuda_graph": use_graph, "ms_per_step_python_launched": ms_launch / args.steps,
        "launch_note": "gpu_launches counted in the Python-launched pass; the graphed pass replays the same kernels",
        "roofline": {"bound": "tensor", "kernel": "conv3x3_tc2_kernel (fwd + dgrad) and wgrad3x3_tc_kernel",
                     "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                     "peak_source": f"{peak_src} bf16_tflops_sustained", "traffic": None,
                     "kernel_ms_per_step": tc_ms / args.steps, "share_of_step": tc_ms / ms},
        "kernels": per_kernel,
        "cpu_baseline": cpu_train,
        "source_train": src,
        "joint_fixmatch": joint,
        "model_tflops": TRAIN_FLOP_PER_PX * float(Bt) * HW * HW * world * args.steps / (ms * 1e-3) / 1e12,
    }


def run_ours(args):
    import torch
    import torch.distributed as dist
    from probabilistic_domain_adaptation_b200 import _lib, consensus, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.mode == "train":  # training leg only (profiling aid): same object, promoted to the top level
        train = run_train_leg(args, dev, world, rank, local, lib, barrier)
        if rank == 0:
            line = dict(train)
            line.update({"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
                         "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic"})
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return

    T, HW, S = args.tiles, args.size, args.samples
    model = make_model(dev).eval()
    g = torch.Generator().manual_seed(1 + rank)
    host_x = torch.randn(T, 1, HW, HW, generator=g).pin_memory()
    eps = torch.randn(S, T, 6, generator=torch.Generator().manual_seed(3)).to(dev)
    x_dev = host_x.to(dev)
    out_mean = torch.empty(T, 1, HW, HW, dtype=torch.float32).pin_memory()
    out_mask = torch.empty(T, 1, HW, HW, dtype=torch.int64).pin_memory()

    def step_resident():
        return consensus.sample_from_teacher(model, x_dev, S, do_consensus_masking=True, eps=eps)

    predictor = consensus.HostPredictor(model, S, True)
    # two sets of pinned output buffers: the device->host copy of step i overlaps the compute of step i+1
    outs = [(out_mean, out_mask),
            (torch.empty_like(out_mean).pin_memory(), torch.empty_like(out_mask).pin_memory())]
    e2e_i = [0]

    def step_e2e():
        om, oc = outs[e2e_i[0] & 1]
        e2e_i[0] += 1
        return predictor.submit(host_x, om, oc, eps=eps)

    def timed(fn, steps, profile=False):
        barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        ops.PROFILE = [] if profile else None
        lib.pda_reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        predictor.flush()  # the timed region ends when the last device->host copy has landed
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.pda_launch_count()
        prof, ops.PROFILE = ops.PROFILE, None
        clocks = sampler.stop() if sampler else None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, launches, prof, clocks

    for _ in range(args.warmup):
        step_resident()
    ms, launches, prof, clocks = timed(step_resident, args.steps, profile=True)
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    ms_e2e, _, _, _ = timed(step_e2e, args.steps)

    px_samples = float(T) * HW * HW * S * world
    value = px_samples * args.steps / (ms * 1e-3)
    e2e_value = px_samples * args.steps / (ms_e2e * 1e-3)

    # roofline of the dominant kernel (tcgen05 conv3x3), timed live per launch on the launching stream
    hbm_peak, tf_peak, peak_src = load_peaks()
    conv_ms = sum(a.elapsed_time(b) for k, a, b, _ in prof if k == "conv3x3_tc")
    conv_flop = sum(w for k, _, _, w in prof if k == "conv3x3_tc")
    n_conv = sum(1 for k, *_ in prof if k == "conv3x3_tc")
    fc_ms = sum(a.elapsed_time(b) for k, a, b, _ in prof if k == "fcomb_mc")
    fc_px = sum(w for k, _, _, w in prof if k == "fcomb_mc")
    achieved = conv_flop / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    fc_gbs = fc_px * 140.0 / (fc_ms * 1e-3) / 1e9 if fc_ms > 0 else 0.0
    fc_tf = fc_px * 2.0 * (4096 + S * 4160) / (fc_ms * 1e-3) / 1e12 if fc_ms > 0 else 0.0

    # BASELINE config 5: S sweep on the same tiles (device-resident timing)
    sweep = None
    if not args.no_extras:
        sweep = {}
        for s_ in (8, 32, 64):
            eps_s = torch.randn(s_, T, 6, generator=torch.Generator().manual_seed(3)).to(dev)
            fn = lambda: consensus.sample_from_teacher(model, x_dev, s_, do_consensus_masking=True, eps=eps_s)  # noqa: E731
            fn()
            ms_s, _, _, _ = timed(fn, max(3, args.steps // 4))
            sweep[str(s_)] = float(T) * HW * HW * s_ * world * max(3, args.steps // 4) / (ms_s * 1e-3)

    # whole-image tiled prediction (blocks of 384 + halo 64 as punet_predictions.py:41-49), image
    # uploaded from pinned host memory, blocks sharded over the ranks
    tiled_res = None
    if not args.no_extras:
        from probabilistic_domain_adaptation_b200 import tiled
        IH = 2048
        himg = (torch.randn(IH, IH, generator=g) * 30 + 100).pin_memory()
        fn = lambda: tiled.predict_with_halo(himg, model, S, (384, 384), (64, 64), 8, rank=rank, world=world)  # noqa: E731
        fn()
        nrep = max(3, args.steps // 4)
        ms_t, _, _, _ = timed(fn, nrep)
        tiled_res = {"metric": "tiled_prediction_image_px_samples_per_s", "value": float(IH) * IH * S * nrep / (ms_t * 1e-3),
                     "unit": "px*samples/s", "ms_per_image": ms_t / nrep,
                     "config": {"workload": f"{IH}x{IH} image, blocks 384 + halo 64 (36 blocks of <= 512x512), S={S}, "
                                            f"per-block standardisation, host image -> device mean-probability image",
                                "parallelism": f"blocks sharded x{world}"}}

    # weak + strong views of one MitoEM-shape batch on the device (every transform applied: p = 1),
    # HBM-bound: statistics pass 4 B/px + per view 4 B/px read + 4 B/px noise + 4 B/px write (noise generation by
    # torch.randn_like is inside the timed region but not in the algorithmic bytes)
    augment_res = None
    if not args.no_extras:
        from probabilistic_domain_adaptation_b200 import augment
        raw_b = (torch.rand(args.train_batch, 1, args.train_size, args.train_size, generator=g) * 255).to(dev)
        aug = augment.DualViewAugmenter(augment.weak_view(1.0), augment.mitoem_strong_view(1.0))
        aug(raw_b)
        nrep = max(5, args.steps)
        ms_a, _, _, _ = timed(lambda: aug(raw_b), nrep)
        px_a = raw_b.numel()
        augment_res = {"metric": "dual_view_augment_img_per_s", "value": args.train_batch * world * nrep / (ms_a * 1e-3),
                       "unit": "img/s", "ms_per_batch": ms_a / nrep,
                       "roofline": {"bound": "hbm", "achieved": px_a * 28.0 * nrep / (ms_a * 1e-3) / 1e9,
                                    "peak": load_peaks()[0], "unit": "GB/s", "algorithmic_bytes_per_px": 28},
                       "config": {"workload": f"{args.train_batch}x1x{args.train_size}x{args.train_size}: statistics + "
                                              "weak view (blur, noise) + strong view (blur, noise, contrast), all applied; "
                                              "decisions sampled on the host per image inside the timed region"}}
        # the training batch is 1 M px (29 MB): launch- and host-sampling-bound.  The kernels alone, on 64 images with
        # pre-sampled decisions (statistics + two views), show the bandwidth they reach
        big = (torch.rand(64, 1, args.train_size, args.train_size, generator=g) * 255).to(dev)
        pw, _, kw = augment.sample_view_params(augment.weak_view(1.0), 64)
        ps, _, ks = augment.sample_view_params(augment.mitoem_strong_view(1.0), 64)
        pw, ps, nz = pw.to(dev), ps.to(dev), torch.randn_like(big)

        def kernels_only():
            st = augment.image_stats(big)
            augment.augment_view(big, pw, kw, noise=nz, stats=st)
            augment.augment_view(big, ps, ks, noise=nz, stats=st)
        kernels_only()
        ms_k, _, _, _ = timed(kernels_only, nrep)
        augment_res["roofline"].update({"achieved": big.numel() * 28.0 * nrep / (ms_k * 1e-3) / 1e9,
                                        "measured_on": f"64x1x{args.train_size}x{args.train_size}, decisions and noise "
                                                       "pre-drawn (3 launches)", "ms": ms_k / nrep})
        augment_res["roofline"]["frac"] = augment_res["roofline"]["achieved"] / augment_res["roofline"]["peak"]
        del raw_b, big, nz

    # BASELINE config 1 shape: ONE 256 x 256 image, S = 16 + consensus mask: latency of a single prediction, launched
    # kernel by kernel from Python and as one CUDA-graph replay (consensus.GraphedMCPredictor)
    small_res = None
    if not args.no_extras:
        xs1 = torch.randn(1, 1, 256, 256, generator=g).to(dev)
        eps1 = torch.randn(S, 1, 6, generator=g).to(dev)
        fn1 = lambda: consensus.sample_from_teacher(model, xs1, S, do_consensus_masking=True, eps=eps1)  # noqa: E731
        fn1()
        nrep = max(20, args.steps)
        ms_1, _, _, _ = timed(fn1, nrep)
        gp = consensus.GraphedMCPredictor(model, xs1, S, do_consensus_masking=True)
        gp(xs1, eps1)
        ms_1g, _, _, _ = timed(lambda: gp(xs1, eps1), nrep)
        small_res = {"metric": "single_image_mc_latency_ms", "python_launched_ms": ms_1 / nrep,
                     "cuda_graph_ms": ms_1g / nrep, "px_samples_per_s_cuda_graph": 256.0 * 256 * S * nrep / (ms_1g * 1e-3),
                     "config": {"workload": f"1x1x256x256 (Lung-XRay shape, BASELINE config 1), S={S}, consensus mask"}}
        del gp

    train = None
    if args.mode in ("train", "both"):
        del x_dev, eps
        torch.cuda.empty_cache()
        train = run_train_leg(args, dev, world, rank, local, lib, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # DRAM traffic per launch: read from the ncu capture of this exact workload (profiles/dram_traffic.json, produced by
    # tools/gpu_traffic.sh + tools/ncu_traffic_summary.py on the build named inside); null for any other workload
    tprof = load_traffic_profile(T, HW, S)
    conv_traffic = tprof["conv3x3_tc"]["dram_bytes_per_launch"] if tprof else None
    conv_traffic_alg = tprof["conv3x3_tc"]["algorithmic_bytes_per_launch"] if tprof else None
    fcomb_traffic = tprof["fcomb_tc"]["dram_bytes_per_launch"] if tprof else None

    eager = None
    if world == 1 and not args.no_cpu_baseline and not args.no_extras:
        eager = eager_gpu_baseline(dev, T, HW, S)
        eager["speedup_vs_eager_fp32_tf32"] = eager["eager_fp32_tf32"]["ms_per_step"] / (ms / args.steps)
        eager["speedup_vs_eager_autocast_bf16"] = eager["eager_autocast_bf16"]["ms_per_step"] / (ms / args.steps)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        csize = min(HW, 512)
        cval, cdt, cores = cpu_reference_run(csize, S, 1, 1, 1)
        cpu = {"value": cval, "unit": "px*samples/s", "cores": cores, "kind": "port",
               "sample": f"1 tile {csize}x{csize}, S={S}, 1 timed step after 1 warm-up ({cdt:.1f} s)"}

    line = {
        "metric": "punet_mc_px_samples_per_s", "value": value, "unit": "px*samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": f"PUNet (64-128-256-512, latent 6) MC inference: forward + S={S} fused "
                               f"Fcomb/sigmoid/mean/consensus-mask on {T} tiles/GPU of 1x{HW}x{HW} "
                               f"(BASELINE config 5 at S={S})",
                   "tiles_per_gpu": T, "tile": HW, "samples": S, "parallelism": f"tile-sharded x{world}",
                   "l2": "activations >> 126 MB L2 (inputs larger than L2, no flush needed)",
                   "arithmetic": "fp16 operands (no-grad path) / bf16 operands (training leg), fp32 accumulation"},
        "e2e": {"value": e2e_value, "unit": "px*samples/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": host_x.numel() * 4,
                "d2h_bytes_per_step": out_mean.numel() * 4 + out_mask.numel() * 8},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv3x3_tc2_kernel (tcgen05 cta_group::2 implicit GEMM, all 31 conv launches of a step)",
                     "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                     "peak_source": f"{peak_src} bf16_tflops_sustained", "traffic": conv_traffic,
                     "traffic_algorithmic": conv_traffic_alg,
                     "traffic_source": (f"profiles/dram_traffic.json ({tprof['build']}; ncu dram__bytes_read+write, mean "
                                        f"per launch)" if tprof else None),
                     "launches": n_conv, "kernel_ms_per_step": conv_ms / args.steps,
                     "share_of_step": conv_ms / ms},
        "roofline_fcomb": {"bound": "hbm (north star) / tensor + CUDA-core epilogue (actual)", "achieved": fc_gbs, "peak": hbm_peak,
                           "unit": "GB/s", "frac": fc_gbs / hbm_peak, "algorithmic_bytes_per_px": 140,
                           "traffic": fcomb_traffic,
                           "tensor_frac": fc_tf / tf_peak,
                           "achieved_tflops": fc_tf, "kernel_ms_per_step": fc_ms / args.steps,
                           "share_of_step": fc_ms / ms},
        "model_tflops": FLOP_PER_PX_FORWARD * T * HW * HW * world * args.steps / (ms * 1e-3) / 1e12,
        "cpu_baseline": cpu,
        "eager_gpu_baseline": eager,
        "train_scaling": (None if train is None else
                          {"metric": train["metric"], "value": train["value"], "unit": train["unit"], "n_gpus": world,
                           "ms_per_step": train["ms_per_step"], "scaling": "weak",
                           "collective": "one NCCL gradient all-reduce per step (109 MB fp32)" if world > 1 else None,
                           "workload": train["config"]["workload"]}),
        "mc_sweep_px_samples_per_s": sweep,
        "tiled_prediction": tiled_res,
        "augment": augment_res,
        "single_image": small_res,
        "train": train,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
