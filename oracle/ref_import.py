"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *unmodified* reference model code from /root/reference (present in the
build container only, never on the GPU box) so that `oracle/make_golden.py` can
pin `oracle/punet_oracle.py` against outputs of the reference itself.

Three interventions are needed:

1. `torch_em.loss.dice.DiceLossWithLogits` (probabilistic_unet.py:8) and
   `matplotlib.pyplot` (my_models/utils.py:3) are absent from this image, so empty
   stand-in modules are registered in `sys.modules` before the import.  The Dice
   stand-in follows torch_em's published algorithm (sigmoid, channel-wise flatten,
   2*sum(p*t) / clamp(sum(p^2)+sum(t^2), eps), 1-score, channel sum).
2. The reference builds block i>0 with C_in = num_filters[i] instead of
   num_filters[i-1] (unet.py:27-28, probabilistic_unet.py:50-51) and crashes on its
   first forward.  The nine affected Conv2d modules are re-created with the intended
   C_in and re-initialised with the reference's own `init_weights`.
3. The reference keeps a module-global `device` (probabilistic_unet.py:15); on CPU
   nothing needs doing.

Everything else (forward / sample / reconstruct / kl_divergence / elbo /
l2_regularisation) is executed verbatim on the reference object.
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("PDA_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "prob_utils", "my_models"))


class _DiceLossWithLogits(nn.Module):
    """Restated from torch_em.loss.dice (unpinned third-party dependency, not vendored):
    channelwise=True, eps=1e-7, reduce_channel='sum'."""

    def forward(self, input_, target):
        p = torch.sigmoid(input_)
        c = p.shape[1]
        pf = p.transpose(0, 1).reshape(c, -1)
        tf = target.transpose(0, 1).reshape(c, -1).to(pf.dtype)
        num = (pf * tf).sum(-1)
        den = (pf * pf).sum(-1) + (tf * tf).sum(-1)
        score = 2.0 * (num / den.clamp(min=1e-7))
        return (1.0 - score).sum()


def _install_stubs():
    if "torch_em" not in sys.modules:
        te = types.ModuleType("torch_em")
        te_loss = types.ModuleType("torch_em.loss")
        te_dice = types.ModuleType("torch_em.loss.dice")
        te_dice.DiceLossWithLogits = _DiceLossWithLogits
        te.loss = te_loss
        te_loss.dice = te_dice
        sys.modules["torch_em"] = te
        sys.modules["torch_em.loss"] = te_loss
        sys.modules["torch_em.loss.dice"] = te_dice
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


def import_reference_modules():
    """Returns (probabilistic_unet module, utils module) of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    # Import the two leaf modules without executing prob_utils/__init__ side effects of
    # other sub-packages (trainers need the real torch_em).
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # our repo ships a drop-in `prob_utils` shim; make sure the reference one wins here
    for k in [k for k in sys.modules if k == "prob_utils" or k.startswith("prob_utils.")]:
        mod = sys.modules[k]
        f = getattr(mod, "__file__", "") or ""
        if not f.startswith(REFERENCE_ROOT):
            del sys.modules[k]
    pkg = types.ModuleType("prob_utils")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "prob_utils")]
    sys.modules["prob_utils"] = pkg
    sub = types.ModuleType("prob_utils.my_models")
    sub.__path__ = [os.path.join(REFERENCE_ROOT, "prob_utils", "my_models")]
    sys.modules["prob_utils.my_models"] = sub
    pu = importlib.import_module("prob_utils.my_models.probabilistic_unet")
    ut = importlib.import_module("prob_utils.my_models.utils")
    return pu, ut


def make_reference_model(num_filters=(64, 128, 256, 512), latent_dim=6, no_convs_fcomb=3, beta=1.0,
                         consensus_masking=False, rl_swap=False, seed=0):
    """Reference ProbabilisticUnet with the intended channel wiring (see module docstring)."""
    pu, ut = import_reference_modules()
    nf = list(num_filters)
    torch.manual_seed(seed)
    m = pu.ProbabilisticUnet(1, 1, nf, latent_dim, no_convs_fcomb, beta, consensus_masking, rl_swap)
    for i in range(1, len(nf)):
        conv = nn.Conv2d(nf[i - 1], nf[i], 3, stride=1, padding=1)
        ut.init_weights(conv)
        m.unet.contracting_path[i].layers[1] = conv
    for enc in (m.prior.encoder, m.posterior.encoder):
        # the conv that follows each AvgPool2d
        idx = [j + 1 for j, l in enumerate(enc.layers) if isinstance(l, nn.AvgPool2d)]
        for i, j in enumerate(idx, start=1):
            conv = nn.Conv2d(nf[i - 1], nf[i], 3, padding=1)
            ut.init_weights(conv)
            enc.layers[j] = conv
    return m


def restore_repo_prob_utils():
    """Drop the reference's prob_utils from sys.modules so the repo's shim can be imported again."""
    for k in [k for k in sys.modules if k == "prob_utils" or k.startswith("prob_utils.")]:
        del sys.modules[k]
    if REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)
