"""Differentiable (training) variants of the ops: torch.autograd.Function wrappers over the backward kernels.

These replace what torch.autograd runs behind `loss.backward()` in the reference's step bodies
(punet_trainer.py:24-36, mean_teacher_trainer.py:111-119, ...: cuDNN dgrad / wgrad and ATen elementwise
backward).  Activations and their gradients are NHWC bf16; parameter gradients are fp32 in the parameter's
own layout (OIHW), so torch optimizers / GradScaler work unchanged.  autograd itself (graph, accumulation
into .grad) is the plumbing; every tensor-sized arithmetic step is a libpda_b200 kernel.
"""
import os

import torch

from . import ops
from .autograd_ops import packed_weight


class Conv3x3Fn(torch.autograd.Function):
    """conv3x3(pad 1) over the channel concat (x, src1) + bias + ReLU, optional fused 2x2 average pool output.
    backward: dZ = (dFull + pool^T dPool) * (Y > 0); dgrad = the same tcgen05 conv on dZ with the rotated
    weights (one launch per concat segment, contiguous row slices of the packed operand); wgrad on tensor cores."""

    @staticmethod
    def forward(ctx, x, src1, weight, bias, conv, relu, want_pool):
        full, pool = ops.conv3x3(x, src1, packed_weight(conv), bias.detach(), relu, True, want_pool)
        ctx.conv, ctx.relu, ctx.want_pool = conv, relu, want_pool
        ctx.save_for_backward(x, src1, full)
        ctx.set_materialize_grads(False)
        if want_pool:
            return full, pool
        return full, None

    @staticmethod
    def backward(ctx, g_full, g_pool):
        x, src1, full = ctx.saved_tensors
        if g_full is None and g_pool is None:
            return (None,) * 7
        g_full = None if g_full is None else g_full.contiguous()
        g_pool = None if g_pool is None else g_pool.contiguous()
        need_x, need_s1, need_w, need_b = ctx.needs_input_grad[:4]
        dz = ops.relu_pool_bwd(g_full, g_pool, full if ctx.relu else None, shape=full.shape)
        c0 = x.shape[3]
        dx = dsrc1 = dw = db = None
        if need_x or need_s1:
            wrot = packed_weight(ctx.conv, rot180=True)  # (c0 + c1, 9 * cout)
            if need_x:
                dx, _ = ops.conv3x3(dz, None, wrot[:c0], None, relu=False)
            if need_s1 and src1 is not None:
                dsrc1, _ = ops.conv3x3(dz, None, wrot[c0:], None, relu=False)
        if need_w or need_b:
            # the bias gradient (column sums of dZ) rides along in the weight-gradient kernel
            dw, db = ops.conv3x3_wgrad(x, src1, dz, want_bias=True, owner=ctx.conv)
        return dx, dsrc1, dw, db, None, None, None


class ConvStackFn(torch.autograd.Function):
    """One block of the reference nets -- n x [conv3x3 + ReLU] with an optional fused 2x2 average pool of the last
    output (DownConvBlock / the conv_block of UpConvBlock / one Encoder block: unet_blocks.py:16-24,
    probabilistic_unet.py:53-61) -- as a single autograd node, so that the backward can fuse across its layers:

      dZ_last = (dFull + pool^T dPool) * (Y_last > 0)                      one elementwise kernel per block
      for j = last .. 1:  dW_j, db_j = wgrad(Y_{j-1}, dZ_j)                bias gradient from the same kernel
                          dZ_{j-1}   = dgrad(dZ_j, rot180 W_j) * (Y_{j-1} > 0)   ReLU mask fused in the dgrad epilogue
      first layer:        dW_0, db_0 (+ dX, dBridge when the block input needs a gradient)

    args: x (NHWC bf16 or None), src1 (bridge or None), x0 / x1 (fp32 planes of a cin <= 2 first layer or None),
    convs (list of nn.Conv2d containers), pool_last, then weight_0, bias_0, weight_1, bias_1, ...
    """

    @staticmethod
    def forward(ctx, x, src1, x0, x1, convs, pool_last, *params):
        n = len(convs)
        ys = []
        cur = x
        pooled = None
        for j, conv in enumerate(convs):
            last = j == n - 1
            if j == 0 and x0 is not None:
                cur = ops.conv3x3_first(x0, x1, conv.weight.detach(), conv.bias.detach(), True,
                                        dtype=ops.TRAIN_DTYPE)
                if last and pool_last:
                    pooled = ops.avgpool2(cur)
            else:
                cur, pooled = ops.conv3x3(cur, src1 if j == 0 else None, packed_weight(conv), conv.bias.detach(), True,
                                          True, last and pool_last)
            ys.append(cur)
        ctx.convs, ctx.pool_last, ctx.first_planes = convs, pool_last, x0 is not None
        ctx.save_for_backward(x, src1, x0, x1, *ys)
        ctx.set_materialize_grads(False)
        return cur, pooled

    @staticmethod
    def backward(ctx, g_full, g_pool):
        convs = ctx.convs
        n = len(convs)
        x, src1, x0, x1, *ys = ctx.saved_tensors
        none = (None,) * (6 + 2 * n)
        if g_full is None and g_pool is None:
            return none
        g_full = None if g_full is None else g_full.contiguous()
        g_pool = None if g_pool is None else g_pool.contiguous()
        if ctx.first_planes and n == 1 and g_pool is not None:
            # pooled output of a single first-layer conv (not a reference configuration): unfused pool backward
            gp = ops.relu_pool_bwd(None, g_pool, None, shape=ys[0].shape)
            g_full = gp if g_full is None else ops.relu_pool_bwd(g_full, g_pool, None, shape=ys[0].shape)
            g_pool = None
        dz = ops.relu_pool_bwd(g_full, g_pool, ys[-1], shape=ys[-1].shape)
        grads = [None] * (2 * n)
        dx = dsrc1 = None
        for j in range(n - 1, -1, -1):
            conv = convs[j]
            need_w = ctx.needs_input_grad[6 + 2 * j] or ctx.needs_input_grad[7 + 2 * j]
            if j == 0 and ctx.first_planes:
                if need_w:
                    # for n > 1 dz left the dgrad conv of layer 1 already masked by ys[0]; the single-layer block got
                    # its mask from relu_pool_bwd above
                    grads[0], grads[1] = ops.conv3x3_first_bwd(x0, x1, ys[0], dz, premasked=True)
                break
            xin = ys[j - 1] if j > 0 else x
            s1 = src1 if j == 0 else None
            if need_w:
                grads[2 * j], grads[2 * j + 1] = ops.conv3x3_wgrad(xin, s1, dz, want_bias=True, owner=conv)
            if j > 0:
                dz, _ = ops.conv3x3(dz, None, packed_weight(conv, rot180=True), None, relu=False, relu_mask=ys[j - 1])
            else:
                c0 = x.shape[3]
                wrot = packed_weight(conv, rot180=True)  # (c0 + c1, 9 * cout): contiguous row slices per segment
                if ctx.needs_input_grad[0]:
                    dx, _ = ops.conv3x3(dz, None, wrot[:c0], None, relu=False)
                if src1 is not None and ctx.needs_input_grad[1]:
                    dsrc1, _ = ops.conv3x3(dz, None, wrot[c0:], None, relu=False)
        return (dx, dsrc1, None, None, None, None, *grads)


def _train_fmt(t):
    """Inputs that come from a no-grad (fp16) block -- e.g. a frozen encoder -- enter the training path as bf16."""
    return t if t is None or t.dtype == ops.TRAIN_DTYPE else t.to(ops.TRAIN_DTYPE)


def conv_stack_train(convs, x, src1, first_input, pool_last):
    x0, x1 = first_input if first_input is not None else (None, None)
    x, src1 = _train_fmt(x), _train_fmt(src1)
    params = []
    for c in convs:
        params += [c.weight, c.bias]
    return ConvStackFn.apply(x, src1, x0, x1, list(convs), pool_last, *params)


def conv3x3_train(x, src1, conv, relu, want_full, want_pool):
    # the full-resolution map is always kept in training: it is the ReLU mask of the backward
    x, src1 = _train_fmt(x), _train_fmt(src1)
    full, pool = Conv3x3Fn.apply(x, src1, conv.weight, conv.bias, conv, relu, want_pool)
    return full, pool


class ConvFirstFn(torch.autograd.Function):
    """First layer (cin 1 or 2 fp32 planes).  The image / label planes never need a gradient."""

    @staticmethod
    def forward(ctx, x0, x1, weight, bias, relu):
        if not relu:
            raise NotImplementedError("first-layer backward assumes the fused ReLU")
        out = ops.conv3x3_first(x0, x1, weight.detach(), bias.detach(), relu, dtype=ops.TRAIN_DTYPE)
        ctx.save_for_backward(x0, x1, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x0, x1, out = ctx.saved_tensors
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            raise NotImplementedError("gradient w.r.t. the input image is not part of the reference's training path")
        dw, db = ops.conv3x3_first_bwd(x0, x1, out, g.contiguous())
        return None, None, dw, db, None


class AvgPool2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.shape = x.shape
        return ops.avgpool2(_train_fmt(x))

    @staticmethod
    def backward(ctx, g):
        return ops.relu_pool_bwd(None, g.contiguous(), None, shape=ctx.shape)


class Upsample2xFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.upsample2x(_train_fmt(x))

    @staticmethod
    def backward(ctx, g):
        return ops.upsample2x_bwd(g.contiguous())


class GaussHeadFn(torch.autograd.Function):
    """Spatial mean + 1x1 conv head of AxisAlignedConvGaussian (probabilistic_unet.py:126-130)."""

    @staticmethod
    def forward(ctx, enc, weight, bias, latent):
        enc = _train_fmt(enc)
        out, scratch = ops.gauss_head_fwd_train(enc, weight.detach(), bias.detach(), latent)
        ctx.latent = latent
        ctx.save_for_backward(enc, weight, scratch)
        return out

    @staticmethod
    def backward(ctx, g):
        enc, weight, scratch = ctx.saved_tensors
        denc, dw, db = ops.gauss_head_bwd(g, weight.detach(), scratch, enc, ctx.latent)
        return denc, dw, db, None


class KlFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mls_q, mls_p):
        ctx.save_for_backward(mls_q, mls_p)
        return ops.kl_diag_gauss(mls_q.contiguous(), mls_p.contiguous())

    @staticmethod
    def backward(ctx, g):
        q, p = ctx.saved_tensors
        dq, dp = ops.kl_diag_gauss_bwd(q.contiguous(), p.contiguous(), g)
        return dq, dp


def kl_op(mls_q, mls_p):
    if torch.is_grad_enabled() and (mls_q.requires_grad or mls_p.requires_grad):
        return KlFn.apply(mls_q, mls_p)
    return ops.kl_diag_gauss(mls_q, mls_p)


class ReconLossFn(torch.autograd.Function):
    """(sum, mean) of BCE-with-logits or Dice-with-logits of (logits * consm, segm * consm)
    (probabilistic_unet.py:347-369), one reduction kernel forward, one elementwise kernel backward."""

    @staticmethod
    def forward(ctx, logits, segm, consm, dice):
        out2, stats = ops.recon_loss_fwd(logits, segm, consm, dice)
        ctx.dice = dice
        ctx.consm = consm
        ctx.save_for_backward(logits, segm, stats)
        return out2

    @staticmethod
    def backward(ctx, g):
        logits, segm, stats = ctx.saved_tensors
        d = ops.recon_loss_bwd(logits, segm, ctx.consm, ctx.dice, stats, g.contiguous().float())
        return d, None, None, None


def recon_loss_op(logits, segm, consm, dice):
    """-> (reconstruction_loss, mean_reconstruction_loss) scalars."""
    logits = logits.contiguous().float()
    segm = segm.detach().contiguous().float()
    if consm is not None:
        consm = consm.detach().contiguous()
        if consm.dtype not in (torch.float32, torch.int64):
            consm = consm.float()
    if torch.is_grad_enabled() and logits.requires_grad:
        out2 = ReconLossFn.apply(logits, segm, consm, dice)
    else:
        out2, _ = ops.recon_loss_fwd(logits, segm, consm, dice)
    return out2[0], out2[1]


class L2NormSumFn(torch.autograd.Function):
    """sum_t ||W_t||_2 over a parameter list (utils.py:32-40): two launches forward, one backward."""

    @staticmethod
    def forward(ctx, cache, *params):
        key = tuple((p.data_ptr(), p.numel()) for p in params)
        if cache.get("key") != key:
            cache["fwd"], cache["bwd"], cache["offsets"], cache["total"] = ops.build_l2_tables(
                [p.detach() for p in params])
            cache["key"] = key
        out, norms = ops.multi_tensor_l2norm_fwd(cache["fwd"], len(params))
        ctx.cache = cache
        ctx.shapes = [p.shape for p in params]
        ctx.params = params
        ctx.save_for_backward(norms)
        return out

    @staticmethod
    def backward(ctx, g):
        (norms,) = ctx.saved_tensors
        c = ctx.cache
        flat = ops.multi_tensor_l2norm_bwd(c["bwd"], norms, g, c["total"])
        grads = [flat[o:o + s.numel()].view(s) for o, s in zip(c["offsets"], ctx.shapes)]
        if DEFER_L2_GRADS:
            # Every regularised parameter also receives a data gradient: returned from here, the engine would sum the
            # two with one ATen add PER PARAMETER (56 launches per step).  The gradients are parked instead and added
            # to param.grad in bulk (one multi-tensor add per gradient bucket / per backward): same .grad afterwards.
            _defer_reg_grads(ctx.params, grads)
            return (None,) * (1 + len(grads))
        return (None, *grads)


# ---- regulariser gradients added in bulk --------------------------------------------------------------------------
# (PDA_DEFER_L2_GRADS=0 restores gradients returned through the engine -- needed only by callers that differentiate the
# regulariser with torch.autograd.grad instead of .backward(), which never touches param.grad)
DEFER_L2_GRADS = os.environ.get("PDA_DEFER_L2_GRADS", "1") != "0"
_PENDING_REG = {}        # parameter -> its parked regulariser gradient (a view of one flat buffer per module)
_LAUNCHED = []           # callables (param) -> flat view | None of live GradAllReducers (see parallel.py)


def _defer_reg_grads(params, grads):
    first = not _PENDING_REG
    for p, g in zip(params, grads):
        if p.requires_grad:
            prev = _PENDING_REG.get(p)
            _PENDING_REG[p] = g if prev is None else prev + g
    if first and _PENDING_REG:
        torch.autograd.Variable._execution_engine.queue_callback(flush_reg_grads)


def take_reg_grads(params):
    """-> (params that have a parked regulariser gradient, those gradients); the entries are removed.  Called by
    GradAllReducer when a bucket is complete, before it copies / reduces the bucket."""
    ps, gs = [], []
    for p in params:
        g = _PENDING_REG.pop(p, None)
        if g is not None:
            ps.append(p)
            gs.append(g)
    return ps, gs


def flush_reg_grads():
    """End of the backward pass (engine callback): whatever is still parked goes into param.grad with one multi-tensor
    add.  A parameter whose gradient bucket was already handed to a GradAllReducer (the regulariser node ran later than
    the bucket's last data gradient -- not the order the step bodies of this package produce) gets it added to the
    reduced flat gradient instead, after that reduction has completed."""
    if not _PENDING_REG:
        return
    items = list(_PENDING_REG.items())
    _PENDING_REG.clear()
    dst, src = [], []
    for p, g in items:
        late = None
        for probe in _LAUNCHED:
            late = probe(p)
            if late is not None:
                break
        if late is not None:
            dst.append(late)
            src.append(g)
        elif p.grad is None:
            p.grad = g.clone()
        else:
            dst.append(p.grad)
            src.append(g)
    if dst:
        with torch.no_grad():
            torch._foreach_add_(dst, src)


_L2_CACHES = {}   # (storage pointer, numel) of every tensor -> pointer tables; bounded (oldest entry evicted)
_L2_CACHE_LIMIT = 32


def l2_norm_sum(params):
    params = list(params)
    # keyed by what the tables actually contain (pointers and sizes), not by object identity: a recycled id() can
    # never resurrect a stale table, and models that were freed do not pin entries for ever
    ckey = tuple((p.data_ptr(), p.numel()) for p in params)
    cache = _L2_CACHES.get(ckey)
    if cache is None:
        if len(_L2_CACHES) >= _L2_CACHE_LIMIT:
            _L2_CACHES.pop(next(iter(_L2_CACHES)))
        cache = _L2_CACHES[ckey] = {}
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return L2NormSumFn.apply(cache, *params)
    with torch.no_grad():
        return L2NormSumFn.forward(_NoCtx(), cache, *params)


class _NoCtx:
    def save_for_backward(self, *a):
        pass


class FcombTrainFn(torch.autograd.Function):
    """Fcomb for ONE latent sample with a backward (reconstruct() / sample() under autograd).  Forward is the
    fused tensor-core kernel (S = 1); backward recomputes the hidden layers in fp32."""

    @staticmethod
    def forward(ctx, feat, z, w1, b1, w2, b2, w3, b3):
        out = ops.fcomb_mc_consensus(feat, z[None].detach(), w1.detach(), b1.detach(), w2.detach(), b2.detach(),
                                     w3.detach(), b3.detach(), want_mean=False, want_weight=False, want_logits=True)
        ctx.save_for_backward(feat, z, w1, b1, w2, b2, w3)
        ctx.range_flag = out["range_flag"]  # fp16 range flag of the forward: the backward follows the same path
        return out["logits"][0]

    @staticmethod
    def backward(ctx, g):
        feat, z, w1, b1, w2, b2, w3 = ctx.saved_tensors
        dfeat, dw1, db1, dw2, db2, dw3, db3, dz = ops.fcomb_bwd(feat, z.detach(), w1.detach(), b1.detach(),
                                                                w2.detach(), b2.detach(), w3.detach(), g,
                                                                fwd_flag=ctx.range_flag)
        return dfeat, dz, dw1, db1, dw2, db2, dw3, db3


def fcomb_train(feat, z, w, **want):
    """z (S,B,L).  Differentiable only for a single sample with logits as the sole output (the training form)."""
    extras = [k for k in ("want_mean", "want_weight", "want_mask", "want_probs") if want.get(k)]
    if z.shape[0] != 1 or extras or not want.get("want_logits"):
        raise NotImplementedError("autograd through the fused Monte-Carlo kernel: only S=1 logits are differentiable "
                                  "(the reference's consensus sampling runs under torch.no_grad())")
    logits = FcombTrainFn.apply(_train_fmt(feat), z[0], *w)
    return {"mean": None, "weight": None, "mask": None, "logits": logits[None], "probs": None}
