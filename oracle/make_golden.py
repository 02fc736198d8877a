"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt by running the UNMODIFIED reference
model (imported from /root/reference through oracle/ref_import.py) on seeded synthetic inputs.
Runs only in the build container (the reference tree does not travel to the GPU box); the
fixtures it writes are committed and are what pins oracle/punet_oracle.py.

    python oracle/make_golden.py            # rewrites tests/golden/

Fixtures are *derived* from the reference run here -- the reference ships no golden vectors.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import punet_oracle as po  # noqa: E402
from oracle import ref_import  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _sub(t, step=8):
    return t[..., ::step, ::step].contiguous().clone()


def _checksums(t):
    t = t.detach().double()
    return torch.tensor([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()], dtype=torch.float64)


def _load(model, sd):
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return model


def _ref_consensus(model, x, n_samples, seed, do_masking, testing=False):
    """Literal execution of mean_teacher_trainer.py:72-88 (the trainer class itself cannot be imported:
    it needs torch_em) against the reference model object; returns also per-sample logits and z draws."""
    sigmoid = torch.nn.Sigmoid()
    with torch.no_grad():
        model.forward(x, None, training=False)
        torch.manual_seed(seed)
        logits, zs = [], []
        for _ in range(n_samples):
            logits.append(model.sample(testing=testing))
            zs.append(model.z_prior_sample.clone())
        samples = [sigmoid(l) for l in logits]
        consensus = [
            torch.where((s >= 0.9) + (s <= 0.1), torch.tensor(1.), torch.tensor(0.)) for s in samples
        ]
        y = torch.stack(samples, dim=0).sum(dim=0) / n_samples
        z = torch.stack(consensus, dim=0).sum(dim=0) / n_samples
        if do_masking:
            z = torch.where(z == 1, 1, 0)
    return y, z, torch.stack(logits, 0), torch.stack(zs, 0)


def mc_case(name, b, h, w, s, gain):
    sd = po.make_state_dict(seed=0, last_layer_gain=gain)
    model = _load(ref_import.make_reference_model(), sd).eval()
    x = torch.randn(b, 1, h, w, generator=torch.Generator().manual_seed(1))
    torch.manual_seed(3)
    eps = torch.stack([torch.randn(b, 6) for _ in range(s)], 0)
    y, zw, logits, zs = _ref_consensus(model, x, s, 3, False)
    _, zm, logits2, _ = _ref_consensus(model, x, s, 3, True)
    assert torch.equal(logits, logits2)
    # prediction path uses .sample(testing=True) (punet_predictions.py:31): same values for the same seed
    y_t, _, logits_t, zs_t = _ref_consensus(model, x, s, 3, False, testing=True)
    mu = model.prior_latent_space.base_dist.loc
    sigma = model.prior_latent_space.base_dist.scale
    assert torch.allclose(zs, mu[None] + sigma[None] * eps, atol=0, rtol=0), "rsample != mu + sigma*randn"
    assert torch.allclose(zs_t, zs, atol=1e-6), "sample(testing=True) draws differ from rsample draws"
    feat = model.unet_features
    out = {
        "desc": f"reference MC sampling + consensus, x=randn(seed1) ({b},1,{h},{w}), S={s}, weights=make_state_dict(0, gain={gain})",
        "b": b, "h": h, "w": w, "s": s, "gain": gain,
        "x": x, "eps": eps, "z_draws": zs,
        "mu_p": mu.clone(), "log_sigma_p": torch.log(sigma).clone(),
        "feat_sub": _sub(feat), "feat_checksum": _checksums(feat),
        "logits": logits if h * w <= 4096 else _sub(logits, 4),
        "logits_checksum": _checksums(logits),
        "y": y, "z_weight": zw, "z_mask": zm.to(torch.uint8),
        "mask_fraction": zm.float().mean().item(),
        "weights_checksum": _checksums(torch.cat([v.flatten() for v in sd.values()])),
    }
    torch.save(out, os.path.join(GOLD, f"{name}.pt"))
    print(name, "mask fraction", out["mask_fraction"], "logit range", logits.min().item(), logits.max().item())


def train_case(name, b, h, w, rl_swap, consm_kind, with_grads):
    sd = po.make_state_dict(seed=0, last_layer_gain=4.0)
    model = _load(ref_import.make_reference_model(consensus_masking=consm_kind is not None, rl_swap=rl_swap), sd)
    model.train()
    _, ut = ref_import.import_reference_modules()
    x = torch.randn(b, 1, h, w, generator=torch.Generator().manual_seed(1))
    if consm_kind is None:
        segm = (torch.rand(b, 1, h, w, generator=torch.Generator().manual_seed(2)) > 0.5).float()
        consm = None
    else:
        # soft pseudo-label + consensus as a teacher would produce them
        segm = torch.rand(b, 1, h, w, generator=torch.Generator().manual_seed(2))
        k = torch.randint(0, 17, (b, 1, h, w), generator=torch.Generator().manual_seed(5))
        consm = k.float() / 16 if consm_kind == "weight" else torch.where(k >= 12, 1, 0)
    eps_post = torch.randn(b, 6, generator=torch.Generator().manual_seed(4))
    model.forward(x, segm, training=True)
    torch.manual_seed(4)
    elbo = model.elbo(segm, consm)
    zq = model.posterior_latent_space.base_dist
    assert torch.equal(torch.randn(b, 6, generator=torch.Generator().manual_seed(4)), eps_post)
    # step body punet_trainer.py:31-34
    reg = ut.l2_regularisation(model.posterior) + ut.l2_regularisation(model.prior) + \
        ut.l2_regularisation(model.fcomb.layers)
    loss = -elbo + 1e-5 * reg
    out = {
        "desc": f"reference forward(training=True)+elbo, ({b},1,{h},{w}), rl_swap={rl_swap}, consm={consm_kind}",
        "b": b, "h": h, "w": w, "rl_swap": rl_swap, "consm_kind": consm_kind,
        "x": x, "segm": segm, "consm": consm, "eps_post": eps_post,
        "elbo": elbo.detach().clone(), "kl": model.kl.detach().clone(),
        "reconstruction_loss": model.reconstruction_loss.detach().clone(),
        "mean_reconstruction_loss": model.mean_reconstruction_loss.detach().clone(),
        "reconstruction": model.reconstruction.detach().clone(),
        "mu_q": zq.loc.detach().clone(), "log_sigma_q": torch.log(zq.scale).detach().clone(),
        "mu_p": model.prior_latent_space.base_dist.loc.detach().clone(),
        "log_sigma_p": torch.log(model.prior_latent_space.base_dist.scale).detach().clone(),
        "reg": reg.detach().clone(), "loss": loss.detach().clone(),
    }
    if with_grads:
        loss.backward()
        gn, gs = {}, {}
        for k, p in model.named_parameters():
            gn[k] = p.grad.norm().item()
            gs[k] = p.grad.flatten()[:: max(1, p.grad.numel() // 16)][:16].clone()
        out["grad_norms"] = gn
        out["grad_samples"] = gs
    torch.save(out, os.path.join(GOLD, f"{name}.pt"))
    print(name, "elbo", elbo.item(), "kl", model.kl.item(), "reg", reg.item())


def dice_case():
    """prob_utils/my_utils/util.py imports only numpy + torch: the reference function itself is executed."""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location(
        "ref_util", os.path.join(ref_import.REFERENCE_ROOT, "prob_utils", "my_utils", "util.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.RandomState(0)
    seg = rng.rand(2, 96, 80).astype(np.float32)
    gt = (rng.rand(2, 96, 80) > 0.6).astype(np.float32)
    cases = {(ts, tg): ref.dice_score(seg, gt, ts, tg) for ts, tg in [(None, None), (0.5, None), (0.5, 0.5), (0.9, 0.1)]}
    torch.save({"desc": "reference prob_utils/my_utils/util.py:dice_score on rand(seed 0) seg (2,96,80) / gt = rand > 0.6",
                "seg": torch.from_numpy(seg), "gt": torch.from_numpy(gt), "cases": cases},
               os.path.join(GOLD, "dice_score.pt"))
    print("dice_score", cases)


def augment_case():
    """Weak / strong views with the reference's own my_standardize_torch (prob_utils/my_utils/util.py:9-14, executed
    from /root/reference), torchvision's RandomApply / GaussianBlur and the restated torch_em transforms."""
    import importlib.util
    from oracle import augment_oracle as ao
    spec = importlib.util.spec_from_file_location(
        "ref_util", os.path.join(ref_import.REFERENCE_ROOT, "prob_utils", "my_utils", "util.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(11)
    yy, xx = torch.meshgrid(torch.arange(72.0), torch.arange(104.0), indexing="ij")
    base = torch.stack([100 + 60 * torch.sin(xx / (5 + 2 * i)) * torch.cos(yy / (9 - i)) for i in range(3)])[:, None]
    raw = base + 25.0 * torch.rand(3, 1, 72, 104, generator=g)   # structured image + pixel noise, uint8-like range
    cases = {}
    for name, (wk, sk, seed) in {"scripts_seed5": (ao.WEAK, ao.STRONG, 5), "scripts_seed8": (ao.WEAK, ao.STRONG, 8),
                                 "all_on_seed3": (ao.ALL_ON, ao.ALL_ON, 3)}.items():
        v1, v2 = ao.dual_views(raw, wk, sk, seed, standardize=ref.my_standardize_torch)
        cases[name] = {"weak_kw": wk, "strong_kw": sk, "seed": seed, "raw1": v1, "raw2": v2}
        print("augment", name, float(v1.std()), float(v2.std()))
    x = raw[0].clone()
    torch.save({"desc": "dual views of a structured synthetic raw batch (3,1,72,104); my_standardize_torch = the reference's own function",
                "raw": raw, "standardized0": ref.my_standardize_torch(x), "cases": cases},
               os.path.join(GOLD, "augment.pt"))


def main():
    os.makedirs(GOLD, exist_ok=True)
    if "--only-augment" in sys.argv:
        return augment_case()
    dice_case()
    augment_case()
    torch.set_num_threads(os.cpu_count())
    mc_case("mc_64x64_s16", 1, 64, 64, 16, 24.0)
    mc_case("mc_40x72_s4_b2", 2, 40, 72, 4, 24.0)
    mc_case("mc_128x128_s8", 1, 128, 128, 8, 24.0)
    train_case("train_bce_64x64", 2, 64, 64, False, None, True)
    train_case("train_dice_64x64", 2, 64, 64, True, None, False)
    train_case("train_dice_weight_64x64", 2, 64, 64, True, "weight", True)
    train_case("train_dice_mask_48x80", 2, 48, 80, True, "mask", False)
    train_case("train_bce_mask_64x64", 2, 64, 64, False, "mask", False)


if __name__ == "__main__":
    main()
