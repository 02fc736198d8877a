"""CPU: the oracle restatement (oracle/punet_oracle.py) against fixtures produced by the unmodified
reference (oracle/make_golden.py).  Tolerances: fp32 re-association only (F.conv2d both sides)."""
import pytest
import torch

from oracle import punet_oracle as po

MC_CASES = ["mc_64x64_s16", "mc_40x72_s4_b2", "mc_128x128_s8"]
TRAIN_CASES = ["train_bce_64x64", "train_dice_64x64", "train_dice_weight_64x64", "train_dice_mask_48x80",
               "train_bce_mask_64x64"]


def _cs(t):
    t = t.detach().double()
    return torch.tensor([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()], dtype=torch.float64)


@pytest.mark.parametrize("name", MC_CASES)
def test_mc_consensus_matches_reference(golden, name):
    g = golden(name)
    sd = po.make_state_dict(seed=0, last_layer_gain=g["gain"])
    wcs = _cs(torch.cat([v.flatten() for v in sd.values()]))
    assert torch.allclose(wcs, g["weights_checksum"], rtol=1e-12), "synthetic weight recipe drifted"
    with torch.no_grad():
        logits, feat, mu, ls = po.mc_logits(sd, g["x"], g["eps"])
        probs = torch.sigmoid(logits)
        y, zw = po.consensus_from_probs(probs, do_consensus_masking=False)
        _, zm = po.consensus_from_probs(probs, do_consensus_masking=True)
    assert torch.allclose(mu, g["mu_p"], atol=1e-5)
    assert torch.allclose(ls, g["log_sigma_p"], atol=1e-5)
    assert torch.allclose(feat[..., ::8, ::8], g["feat_sub"], atol=1e-4)
    assert torch.allclose(_cs(feat), g["feat_checksum"], rtol=1e-5)
    ref_logits = g["logits"]
    mine = logits if ref_logits.shape == logits.shape else logits[..., ::4, ::4]
    assert torch.allclose(mine, ref_logits, atol=1e-4), (mine - ref_logits).abs().max()
    assert torch.allclose(y, g["y"], atol=1e-5)
    assert zm.dtype == torch.int64
    # masks: bit-exact wherever no sample probability sits within 1e-5 of a threshold
    p = probs
    near = (((p - 0.9).abs() < 1e-5) | ((p - 0.1).abs() < 1e-5)).any(0)
    assert torch.equal(zm[~near], g["z_mask"].long()[~near])
    assert torch.equal(zw[~near], g["z_weight"][~near])
    assert 0.0 < g["mask_fraction"] < 1.0


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_elbo_matches_reference(golden, name):
    g = golden(name)
    sd = po.make_state_dict(seed=0, last_layer_gain=4.0)
    with_grads = "grad_norms" in g
    if with_grads:
        for v in sd.values():
            v.requires_grad_(True)
    ctx = torch.enable_grad() if with_grads else torch.no_grad()
    with ctx:
        out = po.training_loss(sd, g["x"], g["segm"], g["eps_post"], g["consm"], beta=1.0,
                               consensus_masking=g["consm_kind"] is not None, rl_swap=g["rl_swap"])
    for k in ("mu_q", "log_sigma_q", "mu_p", "log_sigma_p"):
        assert torch.allclose(out[k], g[k], atol=1e-5), k
    assert torch.allclose(out["reconstruction"], g["reconstruction"], atol=1e-4)
    for k in ("elbo", "kl", "reconstruction_loss", "mean_reconstruction_loss", "reg", "loss"):
        assert torch.allclose(out[k], g[k], rtol=1e-5, atol=1e-5), (k, out[k].item(), g[k].item())
    if with_grads:
        out["loss"].backward()
        for k, v in sd.items():
            gn = v.grad.norm().item()
            assert abs(gn - g["grad_norms"][k]) <= 1e-3 * abs(g["grad_norms"][k]) + 1e-7, (k, gn, g["grad_norms"][k])
            s = v.grad.flatten()[:: max(1, v.grad.numel() // 16)][:16]
            assert torch.allclose(s, g["grad_samples"][k], rtol=1e-3, atol=1e-6 + 1e-4 * g["grad_norms"][k]), k


def test_state_dict_contract():
    sd = po.make_state_dict(0)
    assert len(sd) == 100
    assert sum(v.numel() for v in sd.values()) == 27_349_593
    assert sd["fcomb.layers.0.weight"].shape == (64, 70, 1, 1)
    assert sd["unet.upsampling_path.0.conv_block.layers.0.weight"].shape == (256, 768, 3, 3)
    assert sd["posterior.encoder.layers.0.weight"].shape == (64, 2, 3, 3)
    assert sd["prior.encoder.layers.21.weight"].shape == (512, 256, 3, 3)


def test_analytic_known_answers():
    mu = torch.randn(3, 6)
    ls = torch.randn(3, 6) * 0.3
    assert torch.allclose(po.kl_analytic(mu, ls, mu, ls), torch.zeros(3), atol=1e-6)
    # Dice of a perfect, saturated prediction -> 0
    t = (torch.rand(2, 1, 16, 16) > 0.5).float()
    assert po.dice_loss_with_logits((t * 2 - 1) * 50, t).abs() < 1e-6
    # consensus in {k/S}; both-sided samples count as consensus (mean_teacher_trainer.py:75-81)
    p = torch.tensor([0.95, 0.05, 0.5, 0.9, 0.1]).view(5, 1, 1, 1, 1)
    y, z = po.consensus_from_probs(p)
    assert abs(z.item() - 4 / 5) < 1e-7
    y, z = po.consensus_from_probs(p[[0, 1, 3, 4]], do_consensus_masking=True)
    assert z.item() == 1 and z.dtype == torch.int64
    assert po.adamt_momentum(0) == 0.0 and po.adamt_momentum(10 ** 6) == 0.999


def test_dice_score_matches_reference(golden):
    g = golden("dice_score")
    for (ts, tg), ref in g["cases"].items():
        assert po.dice_score(g["seg"].numpy(), g["gt"].numpy(), ts, tg) == ref
