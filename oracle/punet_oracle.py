"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch fp32) of the reference's
Probabilistic U-Net path.  Never imported by the product package; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs use it.

Parity status: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF,
imported unmodified (plus the channel-wiring patch, see oracle/ref_import.py) in the
build container; `oracle/make_golden.py` writes those outputs to tests/golden/ and
`tests/test_oracle_golden.py` checks this file against them.  The Dice loss is a
third-party dependency of the reference (torch_em, unpinned, not vendored): its
published algorithm is restated here and in the stub -- that single function is
"parity unpinned".

Every function cites the reference lines it follows (paths relative to /root/reference).
All tensors are NCHW fp32.  Functions are differentiable through torch autograd so the
same file is the backward-pass oracle.
"""
import math

import torch
import torch.nn.functional as F

DEFAULT_FILTERS = (64, 128, 256, 512)


# ----------------------------------------------------------------------------------------
# state_dict layout (100 tensors with the intended channel wiring)
# ----------------------------------------------------------------------------------------
def conv_specs(num_filters=DEFAULT_FILTERS, latent_dim=6, no_convs_fcomb=3, input_channels=1, num_classes=1):
    """Ordered list of (key_prefix, c_out, c_in, k) in the reference's parameter registration order
    (unet, prior, posterior, fcomb: probabilistic_unet.py:251-283)."""
    nf = list(num_filters)
    specs = []
    # unet.py:26-35 contracting path; unet_blocks.py:16-24 (pool occupies index 0 for i>0)
    for i in range(len(nf)):
        cin = input_channels if i == 0 else nf[i - 1]
        base = 0 if i == 0 else 1
        for j in range(3):
            specs.append((f"unet.contracting_path.{i}.layers.{base + 2 * j}", nf[i], cin if j == 0 else nf[i], 3))
    # unet.py:39-43 upsampling path: input = previous output + skip channels
    out = nf[-1]
    for u, i in enumerate(range(len(nf) - 2, -1, -1)):
        cin = out + nf[i]
        out = nf[i]
        for j in range(3):
            specs.append((f"unet.upsampling_path.{u}.conv_block.layers.{2 * j}", out, cin if j == 0 else out, 3))
    # probabilistic_unet.py:44-63 encoders (pool inserted before blocks 1..), :95 1x1 head
    for name, extra in (("prior", 0), ("posterior", num_classes)):
        idx = 0
        for i in range(len(nf)):
            cin = input_channels + extra if i == 0 else nf[i - 1]
            if i != 0:
                idx += 1  # AvgPool2d slot
            for j in range(3):
                specs.append((f"{name}.encoder.layers.{idx}", nf[i], cin if j == 0 else nf[i], 3))
                idx += 2  # conv + relu
        specs.append((f"{name}.conv_layer", 2 * latent_dim, nf[-1], 1))
    # probabilistic_unet.py:168-177 fcomb
    specs.append(("fcomb.layers.0", nf[0], nf[0] + latent_dim, 1))
    for j in range(no_convs_fcomb - 2):
        specs.append((f"fcomb.layers.{2 * (j + 1)}", nf[0], nf[0], 1))
    specs.append(("fcomb.last_layer", num_classes, nf[0], 1))
    return specs


def make_state_dict(seed=0, num_filters=DEFAULT_FILTERS, latent_dim=6, no_convs_fcomb=3, last_layer_gain=1.0):
    """Deterministic synthetic weights with He-normal scale (utils.py:17-22 uses kaiming_normal fan_in/relu);
    parity is always on identical *loaded* weights, never on the init RNG stream."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, cout, cin, k in conv_specs(num_filters, latent_dim, no_convs_fcomb):
        std = math.sqrt(2.0 / (cin * k * k))
        if key.endswith("conv_layer"):
            std = 1.0 / math.sqrt(cin)  # keep log_sigma moderate
        if key.startswith("fcomb"):
            std = 1.0 / math.sqrt(cin)
        w = torch.randn(cout, cin, k, k, generator=g) * std
        b = torch.randn(cout, generator=g) * 0.05
        if key == "fcomb.last_layer":
            w = w * last_layer_gain
        sd[key + ".weight"] = w
        sd[key + ".bias"] = b
    return sd


# ----------------------------------------------------------------------------------------
# network pieces
# ----------------------------------------------------------------------------------------
def _conv_relu(sd, key, x):
    pad = 1 if sd[key + ".weight"].shape[-1] == 3 else 0
    return F.relu(F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], padding=pad))


def _pool(x):
    # unet_blocks.py:17 / probabilistic_unet.py:54: AvgPool2d(2, 2, 0, ceil_mode=True)
    return F.avg_pool2d(x, kernel_size=2, stride=2, padding=0, ceil_mode=True)


def unet_features(sd, x, num_filters=DEFAULT_FILTERS):
    """unet.py:50-69 with apply_last_layer=False; blocks: unet_blocks.py:30-31, :49-59."""
    nlev = len(num_filters)
    skips = []
    for i in range(nlev):
        base = 0 if i == 0 else 1
        if i != 0:
            x = _pool(x)
        for j in range(3):
            x = _conv_relu(sd, f"unet.contracting_path.{i}.layers.{base + 2 * j}", x)
        if i != nlev - 1:
            skips.append(x)
    for u in range(nlev - 1):
        bridge = skips[-u - 1]
        up = F.interpolate(x, mode="bilinear", scale_factor=2, align_corners=True)  # unet_blocks.py:51
        assert up.shape[3] == bridge.shape[3]  # unet_blocks.py:55
        x = torch.cat([up, bridge], 1)  # unet_blocks.py:56 ([up, bridge] order)
        for j in range(3):
            x = _conv_relu(sd, f"unet.upsampling_path.{u}.conv_block.layers.{2 * j}", x)
    return x


def gaussian_params(sd, name, x, segm=None, num_filters=DEFAULT_FILTERS, latent_dim=6):
    """AxisAlignedConvGaussian.forward, probabilistic_unet.py:113-142 -> (mu, log_sigma), each (B, latent)."""
    if segm is not None:
        x = torch.cat((x, segm), dim=1)  # :118
    idx = 0
    for i in range(len(num_filters)):
        if i != 0:
            x = _pool(x)
            idx += 1
        for _ in range(3):
            x = _conv_relu(sd, f"{name}.encoder.layers.{idx}", x)
            idx += 2
    enc = torch.mean(x, dim=2, keepdim=True)  # :126
    enc = torch.mean(enc, dim=3, keepdim=True)  # :127
    mls = F.conv2d(enc, sd[f"{name}.conv_layer.weight"], sd[f"{name}.conv_layer.bias"])  # :130
    mls = mls[:, :, 0, 0]
    return mls[:, :latent_dim], mls[:, latent_dim:]


def fcomb_logits(sd, feat, z, no_convs_fcomb=3):
    """Fcomb.forward, probabilistic_unet.py:200-214: broadcast z over H,W, concat [feat, z], 1x1 convs."""
    b, _, h, w = feat.shape
    zt = z[:, :, None, None].expand(b, z.shape[1], h, w)
    x = torch.cat((feat, zt), dim=1)  # :212
    x = _conv_relu(sd, "fcomb.layers.0", x)
    for j in range(no_convs_fcomb - 2):
        x = _conv_relu(sd, f"fcomb.layers.{2 * (j + 1)}", x)
    return F.conv2d(x, sd["fcomb.last_layer.weight"], sd["fcomb.last_layer.bias"])  # :214 no activation


def kl_analytic(mu_q, ls_q, mu_p, ls_p):
    """kl.kl_divergence(Independent(Normal q), Independent(Normal p)), probabilistic_unet.py:332 -> (B,).
    torch's Normal/Normal rule: 0.5*(var_ratio + t1 - 1 - log(var_ratio)), summed over the event dim."""
    sq, sp = torch.exp(ls_q), torch.exp(ls_p)
    var_ratio = (sq / sp) ** 2
    t1 = ((mu_q - mu_p) / sp) ** 2
    return (0.5 * (var_ratio + t1 - 1.0 - torch.log(var_ratio))).sum(-1)


def dice_loss_with_logits(logits, target, eps=1e-7):
    """torch_em.loss.dice.DiceLossWithLogits (third party, see module docstring); probabilistic_unet.py:347."""
    p = torch.sigmoid(logits)
    c = p.shape[1]
    pf = p.transpose(0, 1).reshape(c, -1)
    tf = target.transpose(0, 1).reshape(c, -1).to(pf.dtype)
    num = (pf * tf).sum(-1)
    den = (pf * pf).sum(-1) + (tf * tf).sum(-1)
    return (1.0 - 2.0 * (num / den.clamp(min=eps))).sum()


def reconstruction_loss(logits, segm, consm=None, consensus_masking=False, rl_swap=False):
    """probabilistic_unet.py:347-369 -> (sum, mean) of the criterion output.
    The mask multiplies logits AND target (:363-364)."""
    if consensus_masking and consm is not None:
        logits = logits * consm
        segm = segm * consm
    if rl_swap:
        r = dice_loss_with_logits(logits, segm)
    else:
        r = F.binary_cross_entropy_with_logits(logits, segm, reduction="none")  # :348 resolves to 'none'
    return r.sum(), r.mean()


def elbo(sd, x, segm, eps_post, consm=None, beta=1.0, consensus_masking=False, rl_swap=False,
         num_filters=DEFAULT_FILTERS, latent_dim=6, no_convs_fcomb=3):
    """forward(training=True) + elbo(), probabilistic_unet.py:285-293, :341-371.
    eps_post: (B, latent) standard-normal draw used by posterior.rsample() (:349)."""
    mu_q, ls_q = gaussian_params(sd, "posterior", x, segm, num_filters, latent_dim)
    mu_p, ls_p = gaussian_params(sd, "prior", x, None, num_filters, latent_dim)
    feat = unet_features(sd, x, num_filters)
    z = mu_q + torch.exp(ls_q) * eps_post
    kl = kl_analytic(mu_q, ls_q, mu_p, ls_p).mean()  # :351-353
    logits = fcomb_logits(sd, feat, z, no_convs_fcomb)  # :356-358
    rsum, rmean = reconstruction_loss(logits, segm, consm, consensus_masking, rl_swap)
    return {
        "elbo": -(rsum + beta * kl),  # :371
        "kl": kl, "reconstruction": logits, "reconstruction_loss": rsum, "mean_reconstruction_loss": rmean,
        "mu_q": mu_q, "log_sigma_q": ls_q, "mu_p": mu_p, "log_sigma_p": ls_p, "z": z, "features": feat,
    }


def l2_regularisation(sd, prefixes=("posterior.", "prior.", "fcomb.layers.")):
    """utils.py:32-40 applied as in punet_trainer.py:32-33: sum over parameter TENSORS of ||W||_2."""
    total = None
    for p in prefixes:
        for k, v in sd.items():
            if k.startswith(p):
                n = v.norm(2)
                total = n if total is None else total + n
    return total


def training_loss(sd, x, segm, eps_post, consm=None, reg_weight=1e-5, **kw):
    """Step body punet_trainer.py:30-34 / mean_teacher_trainer.py:113-117: loss = -elbo + 1e-5 * reg."""
    out = elbo(sd, x, segm, eps_post, consm, **kw)
    out["reg"] = l2_regularisation(sd)
    out["loss"] = -out["elbo"] + reg_weight * out["reg"]
    return out


# ----------------------------------------------------------------------------------------
# Monte-Carlo sampling + consensus (mean_teacher_trainer.py:72-88 and its three copies;
# punet_predictions.py:29-33, :104-124)
# ----------------------------------------------------------------------------------------
def mc_logits(sd, x, eps, num_filters=DEFAULT_FILTERS, latent_dim=6, no_convs_fcomb=3):
    """forward(training=False) then S x sample(): z_s = mu_p + sigma_p * eps[s] (rsample, :302; sample()
    draws torch.normal(mu, sigma), identical values for identical eps).  eps: (S, B, latent).
    Returns logits (S, B, 1, H, W), features, mu_p, log_sigma_p."""
    mu_p, ls_p = gaussian_params(sd, "prior", x, None, num_filters, latent_dim)
    feat = unet_features(sd, x, num_filters)
    out = []
    for s in range(eps.shape[0]):
        z = mu_p + torch.exp(ls_p) * eps[s]
        out.append(fcomb_logits(sd, feat, z, no_convs_fcomb))
    return torch.stack(out, 0), feat, mu_p, ls_p


def consensus_from_probs(probs, upper_thres=0.9, lower_thres=0.1, do_consensus_masking=False):
    """mean_teacher_trainer.py:75-88 on a stack of probabilities (S, B, 1, H, W) fp32.
    Returns (pseudo_label y, consensus z); z is int64 {0,1} when masking, else fp32 k/S."""
    n = probs.shape[0]
    cons = [torch.where((p >= upper_thres) + (p <= lower_thres), torch.tensor(1.0), torch.tensor(0.0))
            for p in probs]
    y = torch.stack(list(probs), dim=0).sum(dim=0) / n
    z = torch.stack(cons, dim=0).sum(dim=0) / n
    if do_consensus_masking:
        z = torch.where(z == 1, 1, 0)
    return y, z


def sample_from_teacher(sd, x, eps, do_consensus_masking=False, **kw):
    """Full restatement of sample_from_teacher / sample_from_weak_model."""
    logits, _, _, _ = mc_logits(sd, x, eps, **kw)
    probs = torch.sigmoid(logits)
    return consensus_from_probs(probs, do_consensus_masking=do_consensus_masking) + (logits,)


def pseudo_labels_from_probs(probs, upper_threshold=0.9, lower_threshold=0.1):
    """punet_predictions.py:113-124 on a stack of probabilities (S, 1, 1, H, W): the mean over the samples and the
    consensus MASK, computed per sample with numpy on the host exactly as there (`>=` / `<=` on float32 arrays against
    Python floats, bool `+` = OR, sum / S == 1).  Returns (mypred float32 (H, W), consensus_mask uint8 (H, W))."""
    import numpy as np
    n = probs.shape[0]
    samples = list(probs)
    mypred = torch.stack(samples, dim=0).sum(dim=0) / n                      # :113
    mypred = mypred.detach().cpu().numpy().squeeze()                         # :114
    masks = []
    for sample in samples:                                                   # :116-120
        sample = sample.detach().cpu().numpy().squeeze()
        masks.append((sample >= upper_threshold) + (sample <= lower_threshold))
    consensus_mask = np.stack(masks, axis=0).sum(axis=0) / n                 # :123
    consensus_mask = np.where(consensus_mask == 1, 1, 0)                     # :124
    return mypred, consensus_mask.astype("uint8")                            # :135


def momentum_update(teacher_sd, student_sd, momentum=0.999):
    """mean_teacher_trainer.py:52-55: t = t * m + p * (1 - m), per tensor, fp32."""
    return {k: teacher_sd[k] * momentum + student_sd[k] * (1.0 - momentum) for k in teacher_sd}


def adamt_momentum(iteration, momentum=0.999):
    """adamt_trainer.py:41 warm-up."""
    return min(1 - 1 / (iteration + 1), momentum)


def dice_score(segmentation, groundtruth, threshold_seg=None, threshold_gt=None):
    """prob_utils/my_utils/util.py:17-44 (numpy): 2 sum(gt * seg) / (sum(gt) + sum(seg) + 1e-7), optional thresholds."""
    import numpy as np
    assert segmentation.shape == groundtruth.shape
    seg = segmentation if threshold_seg is None else segmentation > threshold_seg
    gt = groundtruth if threshold_gt is None else groundtruth > threshold_gt
    nom = 2 * np.sum(gt * seg)
    denom = np.sum(gt) + np.sum(seg)
    return float(nom) / float(denom + 1e-7)


def distribution_alignment(y, source_distribution):
    """fixmatch_trainer.py:77-84."""
    y_binary = torch.where(y >= 0.5, 1, 0)
    _, target = torch.unique(y_binary, return_counts=True)
    target = target / target.sum()
    ratio = source_distribution / target
    return torch.where(y < 0.5, y * ratio[0], y * ratio[1]).clip(0, 1), ratio


# ----------------------------------------------------------------------------------------
# synthetic inputs -- shared by tests and bench
# ----------------------------------------------------------------------------------------
def synthetic_inputs(b, h, w, s=16, latent_dim=6):
    x = torch.randn(b, 1, h, w, generator=torch.Generator().manual_seed(1))
    y = (torch.rand(b, 1, h, w, generator=torch.Generator().manual_seed(2)) > 0.5).float()
    eps = torch.randn(s, b, latent_dim, generator=torch.Generator().manual_seed(3))
    eps_post = torch.randn(b, latent_dim, generator=torch.Generator().manual_seed(4))
    return x, y, eps, eps_post
