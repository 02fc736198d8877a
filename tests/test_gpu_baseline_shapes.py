"""GPU parity AT THE SHAPES THE NUMBERS ARE QUOTED ON (BASELINE.json configs 1 and 5), in RAW units.

Tolerance rule (BASELINE.json north_star, restated once in DESIGN.md section 2): with the same latent draws, sampled
logits are within 1e-2 ABSOLUTE of the reference on the synthetic configuration the benchmark names (seed-0 weights at
the reference's init scale: logits span about +-0.7).  Measured history: with bf16 activations the 21-layer trunk reached
1.2e-2 at 256 x 256 and 1.4e-2 at 1024 x 1024 (max over all logits; 99.99 % below 0.9e-2) -- over the bound.  The no-grad
path therefore stores activations and weights as fp16 (3 more mantissa bits, fp32 accumulation unchanged): 1.7e-3 at
256 x 256.  The error is relative, so for weights whose logits span +-G/2 it scales with G; the tests below print and
assert the raw numbers for gains 1, 8 and 24, split the error into its two stages (conv trunk vs fp16 hidden layer of the
fused Fcomb kernel) and check the consensus mask against the reference's own mask.

The oracle (oracle/punet_oracle.py, fp32 on the host cores) needs ~0.5 s for 256 x 256, ~3 s for 512 x 512 and ~15 s for
1024 x 1024 at S = 16.
"""
import json
import os

import pytest
import torch

from oracle import punet_oracle as po

pytestmark = pytest.mark.gpu

REPORT = {}


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _model(gain):
    from probabilistic_domain_adaptation_b200 import ProbabilisticUnet
    m = ProbabilisticUnet(1, 1, [64, 128, 256, 512], 6, 3, 1.0).to(_dev())
    m.load_state_dict(po.make_state_dict(0, last_layer_gain=gain))
    return m.eval()


def _dump_report():
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_baseline_shapes.json"), "w") as fh:
            json.dump(REPORT, fh, indent=1, sort_keys=True)
    except OSError:
        pass


def _compare(tag, b, h, w, s, gain):
    """Runs oracle and CUDA path on the same weights / input / latent draws; returns the error budget."""
    from probabilistic_domain_adaptation_b200 import ops
    dev = _dev()
    sd = po.make_state_dict(0, last_layer_gain=gain)
    x, _, eps, _ = po.synthetic_inputs(b, h, w, s=s)
    with torch.no_grad():
        ref_logits, ref_feat, mu_p, ls_p = po.mc_logits(sd, x, eps)
        ref_y, ref_mask = po.consensus_from_probs(torch.sigmoid(ref_logits), do_consensus_masking=True)
    m = _model(gain)
    with torch.no_grad():
        m.forward(x.to(dev), None, training=False)
        mean, mask, logits, probs = m.mc_consensus(s, eps=eps.to(dev), do_consensus_masking=True, return_samples=True)
        z = ops.latent_samples(m.prior_latent_space._pda_mls, eps.to(dev))
        w_f = [t.detach() for t in m.fcomb.weights()]
        # stage split: (i) the exact-fp32 Fcomb kernel on the bf16 trunk's features isolates the trunk's contribution;
        # (ii) the tensor-core Fcomb kernel on the ORACLE's features (rounded to bf16 once) isolates the fp16 hidden layer
        l_trunk = ops.fcomb_mc_consensus(m._feat_nhwc, z, *w_f, want_mean=False, want_weight=False, want_logits=True,
                                         precision="fp32")["logits"]
        feat_o = ref_feat.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev)
        z_o = (mu_p[None] + torch.exp(ls_p)[None] * eps).to(dev)
        l_fc_tc = ops.fcomb_mc_consensus(feat_o, z_o, *w_f, want_mean=False, want_weight=False, want_logits=True)["logits"]
        l_fc_32 = ops.fcomb_mc_consensus(feat_o, z_o, *w_f, want_mean=False, want_weight=False, want_logits=True,
                                         precision="fp32")["logits"]
    torch.cuda.synchronize()
    d = (logits.cpu() - ref_logits).abs().flatten()
    feat = m.unet_features.float().cpu()
    mls = m.prior_latent_space._pda_mls.cpu()
    k = max(1, int(d.numel() * 1e-4))
    res = {
        "shape": [b, 1, h, w], "samples": s, "gain": gain,
        "logit_range": [float(ref_logits.min()), float(ref_logits.max())],
        "logit_max_abs_err": float(d.max()),
        "logit_p9999_abs_err": float(d.topk(k).values.min()),
        "logit_mean_abs_err": float(d.mean()),
        "stage_trunk_max_abs_err": float((l_trunk.cpu() - ref_logits).abs().max()),
        "stage_fcomb_fp16_max_abs_err": float((l_fc_tc - l_fc_32).abs().max()),
        "stage_feature_bf16_rounding_max_abs_err": float((l_fc_32.cpu() - ref_logits).abs().max()),
        "feature_max_abs_err": float((feat - ref_feat).abs().max()), "feature_scale": float(ref_feat.abs().max()),
        "mu_max_abs_err": float((mls[:, :6] - mu_p).abs().max()),
        "log_sigma_max_abs_err": float((mls[:, 6:] - ls_p).abs().max()),
        "mean_prob_max_abs_err": float((mean.cpu() - ref_y).abs().max()),
        "mask_agreement": float((mask.cpu() == ref_mask).float().mean()),
        "mask_fraction": float(mask.float().mean()), "mask_fraction_reference": float(ref_mask.float().mean()),
    }
    # consensus bit-exact on identical probabilities (recomputed with the reference's torch ops on the kernel's probs)
    _, cm = po.consensus_from_probs(probs.cpu(), do_consensus_masking=True)
    res["mask_bit_exact_on_own_probs"] = bool(torch.equal(mask.cpu(), cm))
    REPORT[tag] = res
    _dump_report()
    print(tag, json.dumps(res))
    return res


@pytest.mark.parametrize("gain", [1.0, 8.0, 24.0])
def test_config1_256x256_s16_raw_logit_error(gain):
    """BASELINE config 1 exactly: 1 x 1 x 256 x 256, S = 16 + consensus mask.  Gain 1 is the configuration the benchmark
    names (logits +-0.65): the literal 1e-2 bound, measured 1.7e-3.  Gains 8 and 24 stretch the logits to +-5 / +-16 so
    that probabilities saturate and the consensus mask takes both values; the error scales with the gain and still
    stays below 1e-2 * gain / 4."""
    r = _compare(f"c1_256_gain{gain:g}", 1, 256, 256, 16, gain)
    assert r["mask_bit_exact_on_own_probs"]
    assert r["logit_max_abs_err"] < 1e-2 * max(1.0, gain / 4), r   # the literal north-star bound at gain 1
    assert r["logit_p9999_abs_err"] < 0.55e-2 * max(1.0, gain / 4), r   # measured 1.3e-3 / 1.01e-2 / 3.04e-2
    assert r["stage_fcomb_fp16_max_abs_err"] < 0.2e-2 * gain, r          # fp16 A1 and fp16 relu(H2): 0.94e-3 * gain
    assert r["mask_agreement"] > 0.995, r
    if gain >= 24:
        assert 0.0 < r["mask_fraction_reference"] < 1.0, r    # both mask values occur


def test_one_512_tile_s16_raw_logit_error():
    """One 512 x 512 tile (the network input of the tiled prediction driver; a quarter of a config-5 tile)."""
    r = _compare("tile_512_gain8", 1, 512, 512, 16, 8.0)
    assert r["mask_bit_exact_on_own_probs"]
    assert r["logit_max_abs_err"] < 1e-2 * 2.0, r            # measured 1.4e-2 on logits of +-6
    assert r["logit_p9999_abs_err"] < 0.6e-2 * 2.0, r
    assert r["mask_agreement"] > 0.995, r


def test_config5_1024_tile_s16_raw_logit_error():
    """One full 1024 x 1024 tile of BASELINE config 5 against the oracle (~15 s of host time).  (A crop of the 1024 result
    cannot be compared with a 512 result: the bilinear x2 with align_corners=True (unet_blocks.py:51) samples at
    positions that depend on the image size, so the reference itself is not crop-consistent.)"""
    r = _compare("c5_1024_gain8", 1, 1024, 1024, 16, 8.0)
    assert r["mask_bit_exact_on_own_probs"]
    assert r["logit_max_abs_err"] < 1e-2 * 2.0, r            # measured 1.5e-2 on logits of +-6.2
    assert r["logit_p9999_abs_err"] < 0.6e-2 * 2.0, r
    assert r["mask_agreement"] > 0.995, r


def test_config5_1024_unit_gain_literal_tolerance():
    """The literal north-star statement at the bench's tile size: seed-0 weights at the init scale (logits +-0.8),
    1 x 1 x 1024 x 1024, S = 4 (the per-sample error does not depend on S): max |logit error| < 1e-2."""
    r = _compare("c5_1024_gain1_s4", 1, 1024, 1024, 4, 1.0)
    assert r["logit_max_abs_err"] < 1e-2, r
    assert r["logit_p9999_abs_err"] < 0.5e-2, r


def test_config5_batch_of_four_equals_single_tiles():
    """The bench shape itself (4 x 1 x 1024 x 1024, S = 16): size-independent property -- every image of the batch gets
    bit-identical results to the same image processed alone (tiles are independent work units; no cross-image state),
    which ties the full-size run to the single-tile parity test above."""
    from probabilistic_domain_adaptation_b200 import consensus
    dev = _dev()
    m = _model(8.0)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 1, 1024, 1024, generator=g).to(dev)
    eps = torch.randn(16, 4, 6, generator=torch.Generator().manual_seed(3)).to(dev)
    mean4, mask4 = consensus.sample_from_teacher(m, x, 16, do_consensus_masking=True, eps=eps)
    for i in (0, 3):
        mean1, mask1 = consensus.sample_from_teacher(m, x[i:i + 1].contiguous(), 16, do_consensus_masking=True,
                                                     eps=eps[:, i:i + 1].contiguous())
        assert torch.equal(mean4[i:i + 1], mean1) and torch.equal(mask4[i:i + 1], mask1)
    assert 0.0 < float(mask4.float().mean()) < 1.0


def test_fcomb_fp16_range_guard_reroutes_to_fp32():
    """Large-magnitude features / latents (x 1e3 and beyond) leave the packed-fp16 range of the hidden layer: the
    kernel raises its range flag on the device and the same call returns the exact fp32 result (no inf / NaN, no
    silent clipping).  In range, the flag stays down and the tensor-core result is used."""
    from probabilistic_domain_adaptation_b200 import ops
    dev = _dev()
    sd = po.make_state_dict(0, last_layer_gain=1.0)
    k = ["fcomb.layers.0", "fcomb.layers.2", "fcomb.last_layer"]
    w = [sd[f"{n}.{p}"].to(dev).contiguous() for n in k for p in ("weight", "bias")]
    g = torch.Generator().manual_seed(8)
    feat = torch.relu(torch.randn(2, 40, 56, 64, generator=g)).to(torch.bfloat16).to(dev)
    z = torch.randn(4, 2, 6, generator=g).to(dev)
    for scale_f, scale_z, expect_flag in ((1.0, 1.0, 0), (1e3, 1.0, 0), (3e4, 1.0, 1), (1.0, 3e5, 1), (1e5, 1e5, 1)):
        f_s, z_s = (feat.float() * scale_f).to(torch.bfloat16), z * scale_z
        out = ops.fcomb_mc_consensus(f_s, z_s, *w, want_logits=True, want_mask=True, want_weight=False)
        ref = ops.fcomb_mc_consensus(f_s, z_s, *w, want_logits=True, want_mask=True, want_weight=False,
                                     precision="fp32")
        flag = int(out["range_flag"][:1].view(torch.int32).item())
        assert flag == expect_flag, (scale_f, scale_z, flag)
        assert torch.isfinite(out["logits"]).all()
        if expect_flag:
            assert torch.equal(out["logits"], ref["logits"]) and torch.equal(out["mask"], ref["mask"])
            assert torch.equal(out["mean"], ref["mean"])
        else:
            scale = float(ref["logits"].abs().max())
            assert float((out["logits"] - ref["logits"]).abs().max()) < 2e-3 * max(1.0, scale)
    # the hidden layer relu(H2) is rounded to fp16 as well (last layer on the tensor core): a second-layer weight matrix
    # large enough for |H2| to approach the fp16 range must raise the flag too (bound from the row-L1 norm of W2)
    for scale_w2, expect_flag in ((1.0, 0), (3e3, 1)):
        w2 = [w[0], w[1], w[2] * scale_w2, w[3], w[4], w[5]]
        out = ops.fcomb_mc_consensus(feat, z, *w2, want_logits=True, want_mask=True, want_weight=False)
        ref = ops.fcomb_mc_consensus(feat, z, *w2, want_logits=True, want_mask=True, want_weight=False, precision="fp32")
        assert int(out["range_flag"][:1].view(torch.int32).item()) == expect_flag, scale_w2
        if expect_flag:
            assert torch.equal(out["logits"], ref["logits"]) and torch.equal(out["mask"], ref["mask"])


def test_fcomb_two_models_on_two_streams_do_not_interfere():
    """Teacher and student (different weights) launched concurrently on two streams: every call owns its scratch and
    carries its last layer as kernel state, so the results equal the serial ones bit for bit."""
    from probabilistic_domain_adaptation_b200 import ops
    dev = _dev()
    k = ["fcomb.layers.0", "fcomb.layers.2", "fcomb.last_layer"]
    ws = []
    for seed, gain in ((0, 8.0), (1, 3.0)):
        sd = po.make_state_dict(seed, last_layer_gain=gain)
        ws.append([sd[f"{n}.{p}"].to(dev).contiguous() for n in k for p in ("weight", "bias")])
    g = torch.Generator().manual_seed(2)
    feat = torch.relu(torch.randn(2, 256, 256, 64, generator=g)).to(torch.bfloat16).to(dev)
    z = torch.randn(16, 2, 6, generator=g).to(dev)
    serial = [ops.fcomb_mc_consensus(feat, z, *w, want_mask=True, want_weight=False, want_logits=True) for w in ws]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for rep in range(6):
        outs = []
        for st, w in zip(streams, ws):
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                outs.append(ops.fcomb_mc_consensus(feat, z, *w, want_mask=True, want_weight=False, want_logits=True))
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        for o, s_ in zip(outs, serial):
            assert torch.equal(o["logits"], s_["logits"]) and torch.equal(o["mask"], s_["mask"]), rep


def test_punet_pseudo_labels_uint8_mask_matches_reference_arithmetic():
    """consensus.punet_pseudo_labels (punet_predictions.py:104-124): uint8 {0,1} mask + fp32 mean.  Bit-exact against the
    reference's numpy arithmetic applied to the kernel's own probabilities; >= 99 % agreement with the oracle's mask."""
    import numpy as np
    from probabilistic_domain_adaptation_b200 import consensus
    dev = _dev()
    gain, s = 8.0, 8
    m = _model(gain)
    x, _, eps, _ = po.synthetic_inputs(1, 96, 128, s=s)
    mean, mask = consensus.punet_pseudo_labels(m, x.to(dev), prior_samples=s, eps=eps.to(dev))
    assert mask.dtype == torch.uint8 and mean.dtype == torch.float32 and mask.shape == (1, 1, 96, 128)
    with torch.no_grad():
        _, _, _, probs = m.mc_consensus(s, eps=eps.to(dev), testing=True, do_consensus_masking=True,
                                        return_samples=True)
    want_pred, want_mask = po.pseudo_labels_from_probs(probs.cpu())
    assert want_mask.dtype == np.uint8
    assert np.array_equal(mask.cpu().numpy().squeeze(), want_mask)
    assert np.allclose(mean.cpu().numpy().squeeze(), want_pred, atol=1e-6)
    sd = po.make_state_dict(0, last_layer_gain=gain)
    with torch.no_grad():
        ref_logits, _, _, _ = po.mc_logits(sd, x, eps)
    ref_pred, ref_mask = po.pseudo_labels_from_probs(torch.sigmoid(ref_logits))
    assert (mask.cpu().numpy().squeeze() == ref_mask).mean() > 0.99
    assert np.abs(mean.cpu().numpy().squeeze() - ref_pred).max() < 0.05
    assert 0 < int(ref_mask.sum()) < ref_mask.size
