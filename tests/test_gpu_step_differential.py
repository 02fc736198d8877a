"""Whole-step differentials: three consecutive trainer steps on the kernels against the same
three steps of the oracle (fp32 autograd + torch.optim.Adam on the host), for every step body of steps.py:

  mean_teacher_step   mean_teacher_trainer.py:101-131   consensus MASKING (int64), Dice
  fixmatch_step       fixmatch_trainer.py:67-95         consensus WEIGHTING (fp32 k/S) + distribution alignment, BCE
  adamatch_step       adamatch_trainer.py:62-102        joint, consensus weighting, Dice
  adamt_step          adamt_trainer.py:89-128           joint, teacher + warm-up EMA, consensus masking, Dice

Both sides start from identical weights and see identical inputs and identical latent draws: the prior draws are passed
in (`eps`); the posterior draw of elbo() (probabilistic_unet.py:349) comes from the CUDA generator, which is re-seeded
before every step so that the oracle can be given the same numbers.  Compared per step: the loss (relative), the
parameter gradients (cosine over all 100 tensors concatenated), the Adam parameter update (cosine) and, where a
teacher exists, the teacher after the EMA.
"""
import copy
import json
import os

import pytest
import torch

from oracle import punet_oracle as po

pytestmark = pytest.mark.gpu

B, H, W, S, LR, GAIN = 2, 64, 64, 16, 1e-4, 8.0
REPORT = {}


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _structured_labels(b, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(float(h)), torch.arange(float(w)), indexing="ij")
    out = []
    for _ in range(b):
        cy, cx, r = (torch.rand(3, generator=g) * torch.tensor([h, w, h / 3.0]) + torch.tensor([0, 0, 6.0])).tolist()
        out.append(((yy - cy) ** 2 + (xx - cx) ** 2 < r * r).float())
    return torch.stack(out)[:, None]


def _inputs(step):
    g = torch.Generator().manual_seed(50 + step)
    x = torch.randn(B, 1, H, W, generator=g)
    xs = torch.randn(B, 1, H, W, generator=g)
    return {"x1": x + 0.1 * torch.randn(B, 1, H, W, generator=g), "x2": x + 0.25 * torch.randn(B, 1, H, W, generator=g),
            "xs": xs, "ys": _structured_labels(B, H, W, 70 + step),
            "eps": torch.randn(S, B, 6, generator=g)}


def _posterior_draws(dev, seed, n):
    """The next n draws of Normal.rsample() on the CUDA generator after torch.cuda.manual_seed(seed)."""
    torch.cuda.manual_seed(seed)
    return [torch.randn(B, 6, device=dev).cpu() for _ in range(n)]


class OracleTrainer:
    """The reference step bodies on the oracle's functional model (dict of fp32 tensors) with torch.optim.Adam."""

    def __init__(self, sd, kind, consensus_masking, rl_swap, do_masking, source_distribution=None):
        self.student = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        self.teacher = {k: v.clone() for k, v in sd.items()}
        self.opt = torch.optim.Adam(list(self.student.values()), lr=LR)
        self.kind, self.cm, self.rl, self.do_masking, self.src = kind, consensus_masking, rl_swap, do_masking, \
            source_distribution
        self.iteration = 0

    def _loss(self, x, y, eps_post, consm=None):
        return po.training_loss(self.student, x, y, eps_post, consm, beta=1.0, consensus_masking=self.cm,
                                rl_swap=self.rl)["loss"]

    def _pseudo(self, net, x, eps):
        with torch.no_grad():
            y, z, _ = po.sample_from_teacher({k: v.detach() for k, v in net.items()}, x, eps,
                                             do_consensus_masking=self.do_masking)
        return y, z

    def step(self, inp, draws):
        self.opt.zero_grad()
        if self.kind == "mean_teacher":          # mean_teacher_trainer.py:106-131
            y, z = self._pseudo(self.teacher, inp["x1"], inp["eps"])
            loss = self._loss(inp["x2"], y, draws[0], z)
        elif self.kind == "fixmatch":            # fixmatch_trainer.py:72-93
            y, z = self._pseudo(self.student, inp["x1"], inp["eps"])
            if self.src is not None:
                y, _ = po.distribution_alignment(y, self.src)
            loss = self._loss(inp["x2"], y, draws[0], z)
        else:                                    # adamatch_trainer.py:67-96 / adamt_trainer.py:94-120
            sup = self._loss(inp["xs"], inp["ys"], draws[0])
            y, z = self._pseudo(self.teacher if self.kind == "adamt" else self.student, inp["x1"], inp["eps"])
            loss = (sup + self._loss(inp["x2"], y, draws[1], z)) / 2
        loss.backward()
        grads = {k: v.grad.detach().clone() for k, v in self.student.items()}
        before = {k: v.detach().clone() for k, v in self.student.items()}
        self.opt.step()
        delta = {k: self.student[k].detach() - before[k] for k in before}
        if self.kind in ("mean_teacher", "adamt"):
            m = 0.999 if self.kind == "mean_teacher" else po.adamt_momentum(self.iteration, 0.999)
            with torch.no_grad():
                self.teacher = po.momentum_update(self.teacher, {k: v.detach() for k, v in self.student.items()}, m)
        self.iteration += 1
        return float(loss.detach()), grads, delta


def _cos(a, b):
    va = torch.cat([a[k].flatten().double() for k in a])
    vb = torch.cat([b[k].flatten().double() for k in a])
    return float((va * vb).sum() / (va.norm() * vb.norm()).clamp_min(1e-300)), float(va.norm()), float(vb.norm())


CASES = {
    "mean_teacher": dict(consensus_masking=True, rl_swap=True, do_masking=True),
    "fixmatch": dict(consensus_masking=True, rl_swap=False, do_masking=False, source_distribution=[0.7, 0.3]),
    "adamatch": dict(consensus_masking=True, rl_swap=True, do_masking=False),
    "adamt": dict(consensus_masking=True, rl_swap=True, do_masking=True),
}


@pytest.mark.parametrize("kind", list(CASES))
def test_three_steps_match_the_oracle_step(kind):
    from probabilistic_domain_adaptation_b200 import ProbabilisticUnet, consensus, steps
    from probabilistic_domain_adaptation_b200.optim import FusedAdam
    dev = _dev()
    cfg = dict(CASES[kind])
    src = cfg.pop("source_distribution", None)
    sd = po.make_state_dict(0, last_layer_gain=GAIN)
    oracle = OracleTrainer(sd, kind, cfg["consensus_masking"], cfg["rl_swap"], cfg["do_masking"],
                           None if src is None else torch.tensor(src))
    model = ProbabilisticUnet(1, 1, [64, 128, 256, 512], 6, 3, 1.0, consensus_masking=cfg["consensus_masking"],
                              rl_swap=cfg["rl_swap"]).to(dev).train()
    model.load_state_dict(sd)
    teacher = copy.deepcopy(model)
    for p in teacher.parameters():
        p.requires_grad = False
    opt = FusedAdam(model.parameters(), lr=LR)
    ema = consensus.MomentumUpdater(model, teacher)
    names = [n for n, _ in model.named_parameters()]
    rows = []
    for it in range(3):
        inp = _inputs(it)
        gi = {k: v.to(dev) for k, v in inp.items()}
        n_draws = 2 if kind in ("adamatch", "adamt") else 1
        draws = _posterior_draws(dev, 900 + it, n_draws)
        want_loss, want_grads, want_delta = oracle.step(inp, draws)
        before = {n: p.detach().clone() for n, p in model.named_parameters()}
        torch.cuda.manual_seed(900 + it)
        if kind == "mean_teacher":
            loss = steps.mean_teacher_step(model, teacher, opt, ema, gi["x1"], gi["x2"], n_samples=S,
                                           do_consensus_masking=True, eps=gi["eps"])[0]
        elif kind == "fixmatch":
            loss = steps.fixmatch_step(model, opt, gi["x1"], gi["x2"], n_samples=S, do_consensus_masking=False,
                                       source_distribution=src, eps=gi["eps"])[0]
        elif kind == "adamatch":
            loss = steps.adamatch_step(model, opt, gi["xs"], gi["ys"], gi["x1"], gi["x2"], n_samples=S,
                                       do_consensus_masking=False, eps=gi["eps"])[0]
        else:
            loss = steps.adamt_step(model, teacher, opt, ema, it, gi["xs"], gi["ys"], gi["x1"], gi["x2"], n_samples=S,
                                    do_consensus_masking=True, eps=gi["eps"])[0]
        got_loss = float(loss)
        got_grads = {n: p.grad.detach().cpu() for n, p in model.named_parameters()}
        got_delta = {n: (p.detach() - before[n]).cpu() for n, p in model.named_parameters()}
        gcos, gn_ref, gn = _cos(want_grads, got_grads)
        dcos, _, _ = _cos(want_delta, got_delta)
        row = {"step": it, "loss": got_loss, "loss_oracle": want_loss,
               "loss_rel_err": abs(got_loss - want_loss) / abs(want_loss), "grad_cosine": gcos,
               "grad_norm_ratio": gn / gn_ref, "adam_delta_cosine": dcos}
        if kind in ("mean_teacher", "adamt"):
            terr = max(float((t.detach().cpu() - oracle.teacher[n]).abs().max())
                       for n, t in zip(names, teacher.parameters()))
            row["teacher_max_abs_diff"] = terr
        rows.append(row)
        print(kind, json.dumps(row))
    REPORT[kind] = rows
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, f"step_differential_{kind}.json"), "w") as fh:
            json.dump(rows, fh, indent=1)
    except OSError:
        pass
    for row in rows:
        # step 0 starts from identical weights: the loss differs only by the bf16 arithmetic of the student forward and
        # by pseudo-label pixels that sit at a threshold.  Later steps also carry the difference of the earlier updates
        # (Adam's first updates are lr * g / (|g| + eps): sign-like, so gradient elements near zero flip sign).
        assert row["loss_rel_err"] < (5e-3 if row["step"] == 0 else 2e-2), row
        assert row["grad_cosine"] > (0.999 if row["step"] == 0 else 0.995), row
        assert abs(row["grad_norm_ratio"] - 1.0) < 0.05, row
        assert row["adam_delta_cosine"] > 0.93, row
        if "teacher_max_abs_diff" in row:
            assert row["teacher_max_abs_diff"] < 8 * LR, row
