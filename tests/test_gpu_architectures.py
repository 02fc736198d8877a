"""Architectures other than the scripts' 64-128-256-512 / no_convs_fcomb=3: the reference's DEFAULT constructor
(probabilistic_unet.py:227-237: num_filters=[32, 64, 128, 192], no_convs_fcomb=4, beta=10) and widths that are not
multiples of 64.  Widths run zero-padded to the 64-channel tiles of the conv kernels (results of the real channels are
those of the unpadded net); no_convs_fcomb != 3 runs forward / Monte-Carlo inference on the general-depth fp32 kernel."""
import pytest
import torch

from oracle import punet_oracle as po

NF = (32, 64, 128, 192)


def test_default_constructor_builds_on_cpu_with_the_reference_state_dict_layout():
    from probabilistic_domain_adaptation_b200 import ProbabilisticUnet
    m = ProbabilisticUnet()                                   # reference defaults
    want = po.make_state_dict(0, num_filters=NF, no_convs_fcomb=4)
    got = m.state_dict()
    assert list(got.keys()) == list(want.keys())
    assert all(got[k].shape == want[k].shape for k in want)
    assert m.beta == 10.0 and m.no_convs_fcomb == 4 and m.num_filters == list(NF)
    with pytest.raises(NotImplementedError):
        ProbabilisticUnet(1, 1, [96, 128, 256, 512], 6, 3, 1.0)   # the fused Fcomb kernels are 64 wide


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


@pytest.mark.gpu
def test_default_constructor_architecture_predicts_like_the_reference():
    from probabilistic_domain_adaptation_b200 import ProbabilisticUnet, consensus
    dev = _dev()
    gain = 4.0
    sd = po.make_state_dict(0, num_filters=NF, no_convs_fcomb=4, last_layer_gain=gain)
    m = ProbabilisticUnet().to(dev).eval()
    m.load_state_dict(sd)
    x, _, eps, _ = po.synthetic_inputs(2, 48, 72, s=6)
    with torch.no_grad():
        ref, feat, mu, ls = po.mc_logits(sd, x, eps, num_filters=NF, no_convs_fcomb=4)
        m.forward(x.to(dev), None, training=False)
        mean, mask, logits, probs = m.mc_consensus(6, eps=eps.to(dev), do_consensus_masking=True, return_samples=True)
        torch.manual_seed(3)
        one = m.sample(testing=True)
    assert m.unet_features.shape == (2, 32, 48, 72)
    fe = (m.unet_features.float().cpu() - feat).abs().max().item()
    assert fe < 5e-3 * feat.abs().max().item(), fe
    err = (logits.cpu() - ref).abs().max().item()
    print("default architecture: max |logit err| =", err, "range", ref.abs().max().item())
    assert err < 1e-2, err
    _, cm = po.consensus_from_probs(probs.cpu(), do_consensus_masking=True)
    assert torch.equal(mask.cpu(), cm) and one.shape == (2, 1, 48, 72)
    y, z = consensus.sample_from_teacher(m, x.to(dev), 6, do_consensus_masking=False, eps=eps.to(dev))
    assert torch.allclose(y, mean) and z.dtype == torch.float32
    # training through a 4-layer Fcomb is not implemented: a clear error, not a wrong gradient
    m.train()
    m.forward(x.to(dev), torch.zeros_like(x).to(dev), training=True)
    with pytest.raises(NotImplementedError):
        m.elbo(torch.zeros_like(x).to(dev))


@pytest.mark.gpu
@pytest.mark.parametrize("rl_swap", [True, False])
def test_padded_widths_train_like_the_reference(rl_swap):
    """num_filters = [32, 64, 128, 192] with the scripts' no_convs_fcomb = 3: loss terms and parameter gradients (in the
    REAL parameter shapes) against fp32 autograd through the oracle."""
    from probabilistic_domain_adaptation_b200 import ProbabilisticUnet, steps
    dev = _dev()
    sd = po.make_state_dict(0, num_filters=NF, no_convs_fcomb=3, last_layer_gain=4.0)
    m = ProbabilisticUnet(1, 1, list(NF), 6, 3, 1.0, consensus_masking=True, rl_swap=rl_swap).to(dev).train()
    m.load_state_dict(sd)
    x, _, _, eps_post = po.synthetic_inputs(2, 48, 64)
    yy, xx = torch.meshgrid(torch.arange(48.0), torch.arange(64.0), indexing="ij")
    y = ((yy - 20) ** 2 + (xx - 30) ** 2 < 200).float()[None, None].repeat(2, 1, 1, 1)
    consm = (torch.rand(2, 1, 48, 64, generator=torch.Generator().manual_seed(5)) > 0.3).float()
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    torch.cuda.manual_seed(77)
    eps_post = torch.randn(2, 6, device=dev).cpu()
    out = po.training_loss(ref_sd, x, y, eps_post, consm, beta=1.0, consensus_masking=True, rl_swap=rl_swap,
                           num_filters=NF, no_convs_fcomb=3)
    out["loss"].backward()
    torch.cuda.manual_seed(77)
    loss = steps.punet_loss(m, x.to(dev), y.to(dev), consm.to(dev), use_consm=True)
    loss.backward()
    assert abs(float(loss) - float(out["loss"])) < 5e-3 * abs(float(out["loss"])), (float(loss), float(out["loss"]))
    num = den_a = den_b = 0.0
    for k, p in m.named_parameters():
        assert p.grad is not None and p.grad.shape == ref_sd[k].shape, k
        a, b = p.grad.cpu().double().flatten(), ref_sd[k].grad.double().flatten()
        num += float((a * b).sum()); den_a += float((a * a).sum()); den_b += float((b * b).sum())
    cos = num / (den_a ** 0.5 * den_b ** 0.5)
    print("padded widths: global gradient cosine", cos, "norm ratio", (den_a / den_b) ** 0.5)
    assert cos > 0.995 and abs((den_a / den_b) ** 0.5 - 1.0) < 0.05
