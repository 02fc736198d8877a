"""Op layer between the nn.Module mirror and the C ABI: packed-weight caching and autograd dispatch.

Inference (torch.no_grad) calls go straight to the forward kernels.  The differentiable (training)
variants are torch.autograd.Function wrappers around the dgrad / wgrad / loss-backward kernels.
"""
import torch

from . import ops


# --------------------------------------------------------------------------------------------
# packed bf16 weights, refreshed whenever the fp32 parameter changes (optimizer step, load_state_dict, .to())
# --------------------------------------------------------------------------------------------
def packed_weight(conv, rot180=False, dtype=torch.bfloat16):
    """Cached 16-bit GEMM operand of a conv's weights: bf16 (training path; rot180 = the dgrad operand) or fp16 (no-grad
    path, forward orientation only)."""
    w = conv.weight
    f16 = dtype == torch.float16
    assert not (f16 and rot180)
    key = "_pda_packed_f16" if f16 else ("_pda_packed_rot" if rot180 else "_pda_packed")
    cached = conv.__dict__.get(key)
    stamp = (w._version, w.data_ptr(), w.device)
    if cached is not None and cached[0] == stamp:
        return cached[1]
    packed = ops.pack_conv3x3_weights(w.detach(), rot180=rot180, dtype=dtype)
    conv.__dict__[key] = (stamp, packed)
    return packed


_PACK_CHUNK = 16384


def refresh_packed(module, rot180=True, bf16=True, f16=None):
    """Re-packs the 16-bit operands of every tensor-core conv of `module` in ONE launch (call after an optimizer or
    EMA step; `packed_weight` would otherwise repack lazily, one launch per conv, orientation and format).
    bf16 / rot180: the training operands (forward and dgrad orientation); f16: the no-grad operand -- None = "if this
    module has run a no-grad fp16 forward before" (a student that also labels the weak view, a teacher)."""
    import torch.nn as nn
    from . import _lib
    state = module.__dict__.get("_pda_pack_state")
    convs = state["convs"] if state is not None else None
    if convs is None:  # the module tree of these nets is fixed after construction: walk it once
        convs = [m for m in module.modules() if isinstance(m, nn.Conv2d) and m.kernel_size == (3, 3)
                 and m.in_channels % 64 == 0 and m.out_channels in (64, 128, 256, 512)]
    if not convs or not convs[0].weight.is_cuda:
        return
    if f16 is None:
        f16 = ops.INFER_DTYPE == torch.float16 and any("_pda_packed_f16" in c.__dict__ for c in convs)
    elif f16:
        f16 = ops.INFER_DTYPE == torch.float16
        bf16 = bf16 or not f16          # a no-grad module on the bf16 fallback needs the bf16 forward operand
    rot180 = rot180 and bf16
    key = tuple((c.weight.data_ptr(), c.weight.device) for c in convs) + (rot180, bf16, f16)
    if state is None or state["key"] != key:
        bufs, rows = [], []
        for c in convs:
            w = c.weight
            cout, cin = w.shape[0], w.shape[1]
            packed = torch.empty((cout, 9 * cin), dtype=torch.bfloat16, device=w.device) if bf16 else None
            rot = torch.empty((cin, 9 * cout), dtype=torch.bfloat16, device=w.device) if rot180 else None
            half = torch.empty((cout, 9 * cin), dtype=torch.float16, device=w.device) if f16 else None
            bufs.append((packed, rot, half))
            for co0 in range(0, cout, 32):          # one 32 x 32 x 9 tile per block of the pack kernel
                for ci0 in range(0, cin, 32):
                    rows.append((w.data_ptr(), ops._ptr(packed), ops._ptr(rot), ops._ptr(half), cout, cin, co0, ci0))
        table = torch.tensor(rows, dtype=torch.int64).to(convs[0].weight.device)
        state = {"key": key, "bufs": bufs, "table": table, "convs": convs}
        module.__dict__["_pda_pack_state"] = state
    lib = _lib.load()
    _lib.check(lib.pda_pack_conv3x3_weights_multi(state["table"].data_ptr(), state["table"].shape[0],
                                                  ops._stream()), "pack_weights_multi")
    for c, (packed, rot, half) in zip(convs, state["bufs"]):
        w = c.weight
        stamp = (w._version, w.data_ptr(), w.device)
        if packed is not None:
            c.__dict__["_pda_packed"] = (stamp, packed)
        if rot is not None:
            c.__dict__["_pda_packed_rot"] = (stamp, rot)
        if half is not None:
            c.__dict__["_pda_packed_f16"] = (stamp, half)


def invalidate_packed(module):
    """Call after kernels wrote parameter memory behind autograd's back (the fused EMA update)."""
    for m in module.modules():
        m.__dict__.pop("_pda_packed", None)
        m.__dict__.pop("_pda_packed_rot", None)
        m.__dict__.pop("_pda_packed_f16", None)


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
    if src1 is not None and src1.dtype != x.dtype:
        src1 = src1.to(x.dtype)  # (a bridge that was produced under a different grad mode than x)
    return ops.conv3x3(x, src1, packed_weight(conv, dtype=x.dtype), conv.bias.detach(), relu, want_full, want_pool)


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
