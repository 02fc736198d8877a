"""Host-side mirror of prob_utils/my_models/probabilistic_unet.py on the sm_100a kernels.

Same classes, constructor signatures, module tree / state_dict keys, cached attributes and method
contracts as /root/reference/prob_utils/my_models/probabilistic_unet.py:18-371 (intended channel
wiring).  nn.Conv2d children are parameter containers only; the arithmetic
runs in libpda_b200 (no cuDNN, no CPU fallback).

Additions over the reference API (used by the consensus helpers, never required by callers):
  ProbabilisticUnet.mc_consensus(...)  -- S prior samples + sigmoid + mean + consensus in ONE kernel.
"""
import torch
import torch.nn as nn
from torch.distributions import Independent, Normal

from .. import ops
from ..autograd_ops import gauss_head_op
from .unet import Unet
from .unet_blocks import as_nchw, as_nhwc, run_conv_stack, _check_spatial
from .utils import init_weights, init_weights_orthogonal_normal

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


class Encoder(nn.Module):
    """len(num_filters) blocks of no_convs_per_block x [conv3x3 + ReLU], 2x2 average pool between blocks
    (probabilistic_unet.py:18-69)."""

    def __init__(self, input_channels, num_filters, no_convs_per_block, initializers, padding=True, posterior=False,
                 num_classes=None):
        super().__init__()
        self.contracting_path = nn.ModuleList()
        self.input_channels = input_channels
        self.num_filters = list(num_filters)
        if posterior:
            assert num_classes is not None
            self.input_channels += num_classes  # the mask is concatenated on the channel axis (:39-42)
        if not padding:
            raise NotImplementedError("only padding=True is implemented")
        if self.input_channels > 2:
            raise NotImplementedError("first-layer kernel handles 1 (prior) or 2 (posterior) input planes")
        if max(self.num_filters) > 512 or min(self.num_filters) < 1 or self.num_filters[0] > 256:
            raise NotImplementedError("encoder widths must be in 1..512 (first block <= 256); widths that are not "
                                      "multiples of 64 run zero-padded to the 64-channel tiles of the conv kernels")
        mods, self._blocks = [], []
        prev = self.input_channels
        for i, f in enumerate(self.num_filters):
            if i != 0:
                mods.append(nn.AvgPool2d(kernel_size=2, stride=2, padding=0, ceil_mode=True))
            block = []
            for j in range(no_convs_per_block):
                conv = nn.Conv2d(prev if j == 0 else f, f, kernel_size=3, padding=1)
                mods += [conv, nn.ReLU(inplace=True)]
                block.append(len(mods) - 2)
            self._blocks.append(block)
            prev = f
        self.layers = nn.Sequential(*mods)
        self.layers.apply(init_weights)

    def forward_nhwc(self, patch, segm=None):
        """patch (and segm): fp32 (B,1,H,W) planes -> (B, H/2^(L-1), W/2^(L-1), num_filters[-1]) bf16."""
        nblk = len(self._blocks)
        _check_spatial(patch.shape[2], patch.shape[3], nblk)
        x = None
        for i, idxs in enumerate(self._blocks):
            convs = [self.layers[k] for k in idxs]
            last = i == nblk - 1
            # only the pooled map is consumed downstream: the full-resolution map of the block's last conv
            # is never written to HBM
            full, pooled = run_conv_stack(convs, x, first_input=(patch, segm) if i == 0 else None,
                                          pool_last=not last, keep_full=last)
            x = full if last else pooled
        return x

    def forward(self, input):
        planes = input.float()
        segm = planes[:, 1:2].contiguous() if planes.shape[1] == 2 else None
        return as_nchw(self.forward_nhwc(planes[:, 0:1].contiguous(), segm))[:, :self.num_filters[-1]]


class AxisAlignedConvGaussian(nn.Module):
    """Conv net parametrising a diagonal Gaussian over the latent space (probabilistic_unet.py:72-142)."""

    def __init__(self, input_channels, num_filters, no_convs_per_block, latent_dim, initializers, posterior=False,
                 num_classes=None):
        super().__init__()
        self.input_channels = input_channels
        self.channel_axis = 1
        self.num_filters = list(num_filters)
        self.no_convs_per_block = no_convs_per_block
        self.latent_dim = latent_dim
        self.posterior = posterior
        self.name = "Posterior" if posterior else "Prior"
        self.encoder = Encoder(self.input_channels, self.num_filters, self.no_convs_per_block, initializers,
                               posterior=self.posterior, num_classes=num_classes)
        self.conv_layer = nn.Conv2d(self.num_filters[-1], 2 * self.latent_dim, (1, 1), stride=1)
        nn.init.orthogonal_(self.conv_layer.weight, gain=1)
        nn.init.trunc_normal_(self.conv_layer.bias, std=0.001)
        self.mu_log_sigma = None

    def forward_raw(self, input, segm=None):
        """-> (B, 2*latent) fp32 = (mu | log_sigma): encoder, spatial mean, 1x1 conv (:113-137)."""
        if (segm is not None) != self.posterior:
            raise RuntimeError("posterior nets take (patch, segm); prior nets take patch only")
        enc = self.encoder.forward_nhwc(input.float().contiguous(),
                                        None if segm is None else segm.float().contiguous())
        head = self.conv_layer
        if enc.shape[3] != head.in_channels:  # encoder output zero-padded to a multiple of 64 channels
            from .unet_blocks import ConvView
            head = ConvView(head, (head.in_channels,), (enc.shape[3],), head.out_channels)
        self.mu_log_sigma = gauss_head_op(enc, head, self.latent_dim)
        return self.mu_log_sigma

    def forward(self, input, segm=None):
        mls = self.forward_raw(input, segm)
        mu, log_sigma = mls[:, :self.latent_dim], mls[:, self.latent_dim:]
        dist = Independent(Normal(loc=mu, scale=torch.exp(log_sigma), validate_args=False), 1, validate_args=False)
        dist._pda_mls = mls
        return dist


class Fcomb(nn.Module):
    """no_convs_fcomb x conv1x1 combining the U-Net feature map with a latent sample
    (probabilistic_unet.py:145-214).  The kernel never tiles z nor concatenates: see csrc/fcomb.cu."""

    def __init__(self, num_filters, latent_dim, num_output_channels, num_classes, no_convs_fcomb, initializers,
                 use_tile=True):
        super().__init__()
        self.num_channels = num_output_channels
        self.num_classes = num_classes
        self.channel_axis = 1
        self.spatial_axes = [2, 3]
        self.num_filters = list(num_filters)
        self.latent_dim = latent_dim
        self.use_tile = use_tile
        self.no_convs_fcomb = no_convs_fcomb
        self.name = "Fcomb"
        if not use_tile:
            raise NotImplementedError("use_tile=False builds no layers in the reference")
        if no_convs_fcomb < 2 or self.num_filters[0] > 64 or num_classes != 1:
            raise NotImplementedError("Fcomb kernels: no_convs_fcomb >= 2, num_filters[0] <= 64, num_classes = 1 "
                                      "(every reference script: 3, 64, 1; the reference's default constructor: 4, 32, 1)")
        f0 = self.num_filters[0]
        mods = [nn.Conv2d(f0 + latent_dim, f0, kernel_size=1), nn.ReLU(inplace=True)]
        for _ in range(no_convs_fcomb - 2):
            mods += [nn.Conv2d(f0, f0, kernel_size=1), nn.ReLU(inplace=True)]
        self.layers = nn.Sequential(*mods)
        self.last_layer = nn.Conv2d(f0, num_classes, kernel_size=1)
        init = init_weights_orthogonal_normal if initializers["w"] == "orthogonal" else init_weights
        self.layers.apply(init)
        self.last_layer.apply(init)

    def _padded(self, conv, feature_first=False):
        """(weight, bias) of a 1x1 layer zero-padded to the kernels' 64-wide hidden layer (no-op for f0 = 64).  Layer 0's
        input is cat(features, z) (:212): the feature columns are padded to 64, the z columns follow."""
        import torch.nn.functional as F
        f0 = self.num_filters[0]
        w, b = conv.weight, conv.bias
        if f0 == 64:
            return w, b
        if feature_first:
            w = torch.cat((F.pad(w[:, :f0], (0, 0, 0, 0, 0, 64 - f0)), w[:, f0:]), 1)
        elif w.shape[1] == f0:
            w = F.pad(w, (0, 0, 0, 0, 0, 64 - f0))
        if w.shape[0] == f0:
            w, b = F.pad(w, (0, 0, 0, 0, 0, 0, 0, 64 - f0)), F.pad(b, (0, 64 - f0))
        return w, b

    def weights(self):
        """(w1, b1, w2, b2, w3, b3) of the fused kernels (no_convs_fcomb = 3)."""
        l0, l1, ll = self.layers[0], self.layers[2], self.last_layer
        return (*self._padded(l0, True), *self._padded(l1), *self._padded(ll))

    def forward_nhwc(self, feat, z, **want):
        """feat (B,H,W,64) 16-bit (channels beyond num_filters[0] are zero), z (S,B,L) -> dict of outputs."""
        from ..autograd_ops import _needs_grad
        if self.no_convs_fcomb != 3:
            return self._forward_deep(feat, z, **want)
        w = self.weights()
        if _needs_grad(feat, z, *w):
            from ..training import fcomb_train
            return fcomb_train(feat, z, w, **want)
        return ops.fcomb_mc_consensus(feat, z, *[t.detach() for t in w], **want)

    def _forward_deep(self, feat, z, **want):
        """no_convs_fcomb != 3 (the reference's default constructor builds 4): general-depth fp32 kernel, forward only."""
        from ..autograd_ops import _needs_grad
        convs = [m for m in self.layers if isinstance(m, nn.Conv2d)]
        if _needs_grad(feat, z, *[p for c in convs for p in (c.weight, c.bias)], self.last_layer.weight):
            raise NotImplementedError("training through Fcomb is implemented for no_convs_fcomb=3 (every reference "
                                      "script); other depths run forward / Monte-Carlo inference only")
        w1, b1 = self._padded(convs[0], True)
        mids = [self._padded(c) for c in convs[1:]]
        wmid = torch.stack([w.reshape(64, 64) for w, _ in mids]).contiguous() if mids else None
        bmid = torch.stack([b for _, b in mids]).contiguous() if mids else None
        w3, b3 = self._padded(self.last_layer)
        return ops.fcomb_mc_consensus_deep(feat, z, w1.detach(), b1.detach(), wmid, bmid, w3.detach(), b3.detach(),
                                           **want)

    def forward(self, feature_map, z):
        from .unet_blocks import pad_channels
        out = self.forward_nhwc(pad_channels(as_nhwc(feature_map)), z[None], want_mean=False, want_weight=False,
                                want_logits=True)
        return out["logits"][0]


class ProbabilisticUnet(nn.Module):
    """Probabilistic U-Net (probabilistic_unet.py:217-371): same constructor defaults, stateful
    forward -> sample / reconstruct / kl_divergence / elbo protocol and cached attributes.

    Supported architectures: `num_filters` of any widths with num_filters[0] <= 64 and all <= 512 (widths that are not
    multiples of 64 run zero-padded to the 64-channel tiles of the tensor-core kernels: same results, wasted lanes),
    `num_classes = input_channels = 1`.  `no_convs_fcomb = 3` (every reference script) runs on the fused tensor-core
    kernels, forward and backward; any other depth >= 2 -- the reference's own default is 4 -- runs forward / sampling /
    Monte-Carlo consensus on a plain fp32 kernel and raises NotImplementedError when a gradient through Fcomb is
    requested.  So `ProbabilisticUnet()` with the reference's defaults constructs and predicts; it does not train."""

    def __init__(self, input_channels=1, num_classes=1, num_filters=[32, 64, 128, 192], latent_dim=6,
                 no_convs_fcomb=4, beta=10.0, consensus_masking=False, rl_swap=False):
        super().__init__()
        self.input_channels = input_channels
        self.num_classes = num_classes
        self.num_filters = list(num_filters)
        self.latent_dim = latent_dim
        self.no_convs_per_block = 3
        self.no_convs_fcomb = no_convs_fcomb
        self.initializers = {"w": "he_normal", "b": "normal"}
        self.beta = beta
        self.z_prior_sample = 0
        self.consensus_masking = consensus_masking
        self.rl_swap = rl_swap

        self.unet = Unet(self.input_channels, self.num_classes, self.num_filters, self.initializers,
                         apply_last_layer=False, padding=True).to(device)
        self.prior = AxisAlignedConvGaussian(self.input_channels, self.num_filters, self.no_convs_per_block,
                                             self.latent_dim, self.initializers).to(device)
        self.posterior = AxisAlignedConvGaussian(self.input_channels, self.num_filters, self.no_convs_per_block,
                                                 self.latent_dim, self.initializers, posterior=True,
                                                 num_classes=num_classes).to(device)
        self.fcomb = Fcomb(self.num_filters, self.latent_dim, self.input_channels, self.num_classes,
                           self.no_convs_fcomb, {"w": "orthogonal", "b": "normal"}, use_tile=True).to(device)

    # ------------------------------------------------------------------ reference protocol
    def forward(self, patch, segm, training=True):
        """Caches prior (and posterior, if training) latent spaces and the U-Net features; returns None (:285-293)."""
        if patch.is_cuda and not torch.is_grad_enabled():
            ops.poll_fp16_range(patch.device)  # fp16 range guard of the no-grad path (asynchronous, see ops.py)
        if training:
            self.posterior_latent_space = self.posterior.forward(patch, segm)
        self.prior_latent_space = self.prior.forward(patch)
        self._feat_nhwc = self.unet.forward_nhwc(patch.float().contiguous())
        # logical (B, num_filters[0], H, W) view (channels-last strides; channels padded to 64 inside)
        self.unet_features = as_nchw(self._feat_nhwc)[:, :self.num_filters[0]]

    def sample(self, testing=False):
        """One prior draw (rsample, or sample() when testing) decoded by Fcomb -> logits (B,1,H,W) (:295-309).
        The draw uses the same torch.distributions calls as the reference, hence the same RNG stream."""
        z_prior = self.prior_latent_space.sample() if testing else self.prior_latent_space.rsample()
        self.z_prior_sample = z_prior
        return self._decode(z_prior)

    def _decode(self, z):
        """fcomb(unet_features, z) on the cached NHWC feature map (no layout round trip)."""
        out = self.fcomb.forward_nhwc(self._feat_nhwc, z[None], want_mean=False, want_weight=False, want_logits=True)
        return out["logits"][0]

    def reconstruct(self, use_posterior_mean=False, calculate_posterior=False, z_posterior=None):
        """Decode a posterior sample (:311-322)."""
        if use_posterior_mean:
            z_posterior = self.posterior_latent_space.loc  # AttributeError on Independent, as in the reference
        elif calculate_posterior:
            z_posterior = self.posterior_latent_space.rsample()
        return self._decode(z_posterior)

    def kl_divergence(self, analytic=True, calculate_posterior=False, z_posterior=None):
        """KL(posterior || prior) per batch element (:324-339)."""
        if analytic:
            from ..training import kl_op
            return kl_op(self.posterior_latent_space._pda_mls, self.prior_latent_space._pda_mls)
        if calculate_posterior:
            z_posterior = self.posterior_latent_space.rsample()
        return self.posterior_latent_space.log_prob(z_posterior) - self.prior_latent_space.log_prob(z_posterior)

    def elbo(self, segm, consm=None, analytic_kl=True, reconstruct_posterior_mean=False, view=False):
        """Evidence lower bound: -(reconstruction_loss + beta * KL) (:341-371).  `consm` multiplies logits AND
        target when self.consensus_masking (:363-364)."""
        from ..training import recon_loss_op
        z_posterior = self.posterior_latent_space.rsample()
        self.kl = torch.mean(self.kl_divergence(analytic=analytic_kl, calculate_posterior=False,
                                                z_posterior=z_posterior))
        self.reconstruction = self.reconstruct(use_posterior_mean=reconstruct_posterior_mean,
                                               calculate_posterior=False, z_posterior=z_posterior)
        if view:
            print(self.reconstruction.shape, segm.shape)
        use_mask = self.consensus_masking is True and consm is not None
        self.reconstruction_loss, self.mean_reconstruction_loss = recon_loss_op(
            self.reconstruction, segm, consm if use_mask else None, dice=bool(self.rl_swap))
        return -(self.reconstruction_loss + self.beta * self.kl)

    def release_graph(self):
        """Detaches every tensor the stateful protocol caches on the model (values stay readable: loggers read
        `model.kl`, `model.reconstruction_loss`, ... after the step).  A cached non-leaf tensor keeps the step's whole
        autograd graph alive -- which is what makes `copy.deepcopy` fail after a forward in the reference
         and what ties AccumulateGrad nodes to the stream of an earlier step (CUDA-graph capture)."""
        for name in ("unet_features", "_feat_nhwc", "z_prior_sample", "kl", "reconstruction", "reconstruction_loss",
                     "mean_reconstruction_loss"):
            v = self.__dict__.get(name)
            if torch.is_tensor(v):
                self.__dict__[name] = v.detach()
        for name, net in (("prior_latent_space", self.prior), ("posterior_latent_space", self.posterior)):
            d = self.__dict__.get(name)
            if d is not None and hasattr(d, "_pda_mls"):
                mls = d._pda_mls.detach()
                l = self.latent_dim
                nd = Independent(Normal(loc=mls[:, :l], scale=torch.exp(mls[:, l:]), validate_args=False), 1,
                                 validate_args=False)
                nd._pda_mls = mls
                self.__dict__[name] = nd
            if torch.is_tensor(net.mu_log_sigma):
                net.mu_log_sigma = net.mu_log_sigma.detach()

    # ------------------------------------------------------------------ fused Monte-Carlo path
    @torch.no_grad()
    def mc_consensus(self, n_samples=16, eps=None, z=None, testing=False, upper_thres=0.9, lower_thres=0.1,
                     do_consensus_masking=False, want_consensus=True, return_samples=False):
        """After forward(x, None, training=False): n_samples prior draws -> Fcomb -> sigmoid -> mean and
        consensus, in one kernel.  Equivalent to mean_teacher_trainer.py:74-86 / punet_predictions.py:31-32.

        eps (S,B,L) standard-normal draws or z (S,B,L) latent samples may be supplied; otherwise the draws
        are taken exactly as n_samples successive sample()/rsample() calls would take them (same RNG stream).
        Returns (mean_prob, consensus[, logits, probs]); consensus is int64 {0,1} when do_consensus_masking
        else fp32 k/S; None when want_consensus is False."""
        mls = self.prior_latent_space._pda_mls
        if z is None:
            if eps is None:
                b, l = mls.shape[0], self.latent_dim
                # Normal.rsample/sample draw torch.normal(zeros, ones) of shape (B, L) per call
                eps = torch.stack([torch.randn(b, l, device=mls.device) for _ in range(n_samples)], 0)
            z = ops.latent_samples(mls, eps)
        self.z_prior_sample = z[-1]
        out = self.fcomb.forward_nhwc(self._feat_nhwc, z, upper=upper_thres, lower=lower_thres, want_mean=True,
                                      want_weight=want_consensus and not do_consensus_masking,
                                      want_mask=want_consensus and do_consensus_masking,
                                      want_logits=return_samples, want_probs=return_samples)
        cons = out["mask"] if do_consensus_masking else out["weight"]
        if return_samples:
            return out["mean"], cons, out["logits"], out["probs"]
        return out["mean"], cons
