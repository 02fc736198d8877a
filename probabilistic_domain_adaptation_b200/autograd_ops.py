"""Op layer between the nn.Module mirror and the C ABI: packed-weight caching and autograd dispatch.

Inference (torch.no_grad) calls go straight to the forward kernels.  The differentiable (training)
variants are torch.autograd.Function wrappers around the dgrad / wgrad / loss-backward kernels.
"""
import torch

from . import ops


# --------------------------------------------------------------------------------------------
# packed bf16 weights, refreshed whenever the fp32 parameter changes (optimizer step, load_state_dict, .to())
# --------------------------------------------------------------------------------------------
def packed_weight(conv, rot180=False):
    w = conv.weight
    key = "_pda_packed_rot" if rot180 else "_pda_packed"
    cached = conv.__dict__.get(key)
    stamp = (w._version, w.data_ptr(), w.device)
    if cached is not None and cached[0] == stamp:
        return cached[1]
    packed = ops.pack_conv3x3_weights(w.detach(), rot180=rot180)
    conv.__dict__[key] = (stamp, packed)
    return packed


def invalidate_packed(module):
    """Call after kernels wrote parameter memory behind autograd's back (the fused EMA update)."""
    for m in module.modules():
        m.__dict__.pop("_pda_packed", None)
        m.__dict__.pop("_pda_packed_rot", None)


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def conv3x3_first_op(x0, x1, weight, bias, relu=True):
    if _needs_grad(x0, x1, weight, bias):
        from .training import ConvFirstFn
        return ConvFirstFn.apply(x0, x1, weight, bias, relu)
    return ops.conv3x3_first(x0, x1, weight.detach(), bias.detach(), relu)


def conv3x3_op(x, src1, conv, relu=True, want_full=True, want_pool=False):
    if _needs_grad(x, src1, conv.weight, conv.bias):
        from .training import conv3x3_train
        return conv3x3_train(x, src1, conv, relu, want_full, want_pool)
    return ops.conv3x3(x, src1, packed_weight(conv), conv.bias.detach(), relu, want_full, want_pool)


def avgpool2_op(x):
    if _needs_grad(x):
        from .training import AvgPool2Fn
        return AvgPool2Fn.apply(x)
    return ops.avgpool2(x)


def upsample2x_op(x):
    if _needs_grad(x):
        from .training import Upsample2xFn
        return Upsample2xFn.apply(x)
    return ops.upsample2x(x)


def gauss_head_op(enc, conv_layer, latent):
    if _needs_grad(enc, conv_layer.weight, conv_layer.bias):
        from .training import GaussHeadFn
        return GaussHeadFn.apply(enc, conv_layer.weight, conv_layer.bias, latent)
    return ops.gauss_head(enc, conv_layer.weight.detach(), conv_layer.bias.detach(), latent)


def l2_norm_sum(params):
    from .training import l2_norm_sum as impl
    return impl(params)
