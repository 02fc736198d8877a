"""On-device weak / strong view augmentation.

CPU: the oracle restatement against tests/golden/augment.pt (made with the reference's own my_standardize_torch and
torchvision's RandomApply / GaussianBlur); the host-side parameter sampler consumes the RNG streams exactly like the
reference pipeline.  GPU: the fused kernel against the fixture (same decisions, same noise field)."""
import numpy as np
import pytest
import torch

from oracle import augment_oracle as ao

CASES = ["scripts_seed5", "scripts_seed8", "all_on_seed3"]
TOL = 2e-5   # fp32: separable vs 2-D blur summation order, fp64 vs fp32 statistics


def _spec(kw):
    from probabilistic_domain_adaptation_b200.augment import ViewSpec
    return ViewSpec(blur_p=kw.get("blur_p"), blur_sigma=kw.get("blur_sigma", (0, 5)), noise_p=kw.get("noise_p"),
                    noise_scale=kw.get("noise_scale", (0.0, 0.3)), contrast_p=kw.get("contrast_p"),
                    contrast_alpha=kw.get("contrast_alpha", (0.5, 2)), contrast_mean=kw.get("contrast_mean", 0.0))


def _sample_like_reference(case, raw):
    """Same global-RNG consumption as oracle.dual_views: per sample weak then strong."""
    from probabilistic_domain_adaptation_b200.augment import sample_view_params
    torch.manual_seed(case["seed"])
    np.random.seed(case["seed"])
    rows = [[], []]
    for b in range(raw.shape[0]):
        for v, kw in enumerate((case["weak_kw"], case["strong_kw"])):
            rows[v].append(sample_view_params(_spec(kw), 1, image_shape=tuple(raw.shape[1:]), cpu_noise=True))
    out = []
    for v in range(2):
        params = torch.cat([r[0] for r in rows[v]], 0)
        noise = torch.cat([r[1] if r[1] is not None else torch.zeros((1,) + tuple(raw.shape[1:])) for r in rows[v]], 0)
        out.append((params, noise, max(r[2] for r in rows[v])))
    return out


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(golden, name):
    g = golden("augment")
    case = g["cases"][name]
    v1, v2 = ao.dual_views(g["raw"], case["weak_kw"], case["strong_kw"], case["seed"])
    assert torch.allclose(v1, case["raw1"], atol=1e-6) and torch.allclose(v2, case["raw2"], atol=1e-6)
    assert torch.allclose(ao.my_standardize_torch(g["raw"][0].clone()), g["standardized0"], atol=1e-6)


def _host_model(raw, params, noise):
    """The kernel's arithmetic restated with torch ops on the CPU (checks the sampled decisions, not the kernel)."""
    from torchvision.transforms import functional as TF
    out = []
    for b in range(raw.shape[0]):
        k, sigma, scale, alpha, cmean, nstd = [float(v) for v in params[b, :6]]
        x = raw[b].clone()
        for _ in range(int(nstd)):
            x = ao.my_standardize_torch(x)
        if k > 1:
            x = TF.gaussian_blur(x, [int(k), int(k)], [sigma, sigma])
        x = x + scale * noise[b]
        if alpha != 1.0:
            x = cmean + alpha * (x - cmean)
        out.append(x)
    return torch.stack(out)


@pytest.mark.parametrize("name", CASES)
def test_param_sampler_consumes_rng_like_reference(golden, name):
    g = golden("augment")
    case = g["cases"][name]
    (p1, n1, _), (p2, n2, _) = _sample_like_reference(case, g["raw"])
    assert torch.allclose(_host_model(g["raw"], p1, n1), case["raw1"], atol=TOL)
    assert torch.allclose(_host_model(g["raw"], p2, n2), case["raw2"], atol=TOL)
    if name == "all_on_seed3":
        assert bool((p2[:, 0] > 1).all()) and bool((p2[:, 3] != 1).all())


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_views_match_reference_golden(golden, name):
    from probabilistic_domain_adaptation_b200 import augment
    g = golden("augment")
    case = g["cases"][name]
    raw = g["raw"].cuda()
    stats = augment.image_stats(raw)
    for (params, noise, kmax), want in zip(_sample_like_reference(case, g["raw"]), (case["raw1"], case["raw2"])):
        got = augment.augment_view(raw, params, kmax, noise=noise, stats=stats).cpu()
        assert torch.allclose(got, want, atol=TOL), (got - want).abs().max()


@pytest.mark.gpu
def test_device_views_edge_shapes_and_large_kernels():
    """Ragged sizes (not multiples of the 32-px tile), the largest kernel torch_em can draw (23) and the limit (31),
    blur radius close to the image size, per-image parameters mixed in one batch."""
    from probabilistic_domain_adaptation_b200 import augment
    g = torch.Generator().manual_seed(2)
    for H, W, ks in [(17, 45, 23), (33, 31, 31), (64, 64, 3), (40, 100, 1), (512, 512, 23)]:
        raw = torch.rand(3, 1, H, W, generator=g) * 50 + 3
        params = torch.zeros(3, augment.NPARAM)
        params[:, 0] = torch.tensor([ks, 1, max(ks - 2, 1)])
        params[:, 1] = torch.tensor([2.5, 0.0, 0.7])
        params[:, 2] = torch.tensor([0.0, 0.2, 0.1])
        params[:, 3] = torch.tensor([1.0, 0.5, 2.0])
        params[:, 4] = 0.0
        params[:, 5] = torch.tensor([2, 1, 0])
        noise = torch.randn(3, 1, H, W, generator=g)
        want = _host_model(raw, params, noise)
        got = augment.augment_view(raw.cuda(), params, ks, noise=noise.cuda()).cpu()
        scale = max(1.0, float(want.abs().max()))
        assert torch.allclose(got, want, atol=TOL * scale), ((got - want).abs().max(), H, W, ks)


@pytest.mark.gpu
def test_dual_view_augmenter_contract_and_statistics():
    """(raw, raw1, raw2) like the dual datasets; with every transform off the views are the doubly standardised raw."""
    from probabilistic_domain_adaptation_b200 import augment
    raw = (torch.rand(4, 1, 256, 256) * 200).cuda()
    aug = augment.DualViewAugmenter(augment.ViewSpec(), augment.ViewSpec())
    r, r1, r2 = aug(raw)
    assert r is raw and r1.shape == raw.shape and torch.equal(r1, r2)
    assert torch.allclose(r1.mean(dim=(1, 2, 3)), torch.zeros(4, device="cuda"), atol=1e-4)
    assert torch.allclose(r1.std(dim=(1, 2, 3)), torch.ones(4, device="cuda"), atol=1e-4)
    torch.manual_seed(0)
    np.random.seed(0)
    r, r1, r2 = augment.DualViewAugmenter(augment.weak_view(1.0), augment.livecell_fixmatch_strong_view(1.0))(raw)
    assert torch.isfinite(r1).all() and torch.isfinite(r2).all() and not torch.equal(r1, r2)
    with pytest.raises(Exception):
        augment.augment_view(raw.cpu(), torch.zeros(4, augment.NPARAM), 1)   # no CPU fallback
