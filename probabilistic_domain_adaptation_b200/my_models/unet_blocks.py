"""Host-side mirror of prob_utils/my_models/unet_blocks.py on the sm_100a kernels.

The nn.Conv2d / nn.ReLU / nn.AvgPool2d children exist only as PARAMETER CONTAINERS so that
state_dict keys, parameter order and module paths are identical to the reference
(/root/reference/prob_utils/my_models/unet_blocks.py:7-59); they are never called.  All arithmetic
goes through libpda_b200 (tcgen05 implicit-GEMM conv, fused bias/ReLU/pool, bilinear upsample).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..autograd_ops import conv3x3_first_op, conv3x3_op, avgpool2_op, upsample2x_op
from .utils import init_weights


def as_nhwc(x):
    """Logical NCHW tensor (any dtype / memory format) -> contiguous (B,H,W,C) 16-bit activations (an existing bf16 / fp16
    tensor keeps its format; anything else becomes the training format under autograd, the no-grad format otherwise)."""
    if x.dtype in (torch.bfloat16, torch.float16) and x.dim() == 4 and x.permute(0, 2, 3, 1).is_contiguous():
        return x.permute(0, 2, 3, 1)
    dtype = x.dtype if x.dtype in (torch.bfloat16, torch.float16) else \
        (ops.TRAIN_DTYPE if (torch.is_grad_enabled() and x.requires_grad) else ops.INFER_DTYPE)
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def pad_channels(x_nhwc):
    """Zero-pads the channel dimension of an NHWC tensor to a multiple of 64 (no-op for the reference scripts' widths)."""
    c = x_nhwc.shape[3]
    return x_nhwc if c == pad64(c) else F.pad(x_nhwc, (0, pad64(c) - c)).contiguous()


def as_nchw(x_nhwc):
    """(B,H,W,C) bf16 -> logical (B,C,H,W) view (channels_last strides, no copy)."""
    return x_nhwc.permute(0, 3, 1, 2)


def pad64(c):
    """Channel count the kernels run a width of c at: the next of 64 / 128 / 256 / 512 (the conv tiles need multiples of
    64, the NHWC elementwise kernels address 16-byte channel chunks with shifts: powers of two)."""
    for w in (64, 128, 256, 512):
        if c <= w:
            return w
    raise NotImplementedError(f"channel width {c} > 512")


class ConvView:
    """Stands in for an nn.Conv2d container whose channel counts are not multiples of 64 (the reference's default
    `num_filters=[32, 64, 128, 192]`): `weight` / `bias` are the parameters zero-padded to the 64-channel tiles of the
    tensor-core kernels, laid out segment by segment like the (padded) activations they meet.  The padding is ordinary
    autograd-aware tensor arithmetic, so gradients reach the real parameters through it.  Padded output channels carry
    relu(0) = 0 and meet zero weights in the next layer: the real channels are exactly those of the unpadded net."""

    def __init__(self, conv, segs_real, segs_pad, cout_pad):
        w, b = conv.weight, conv.bias
        parts, off = [], 0
        for r, pd in zip(segs_real, segs_pad):
            seg = w[:, off:off + r]
            parts.append(F.pad(seg, (0, 0, 0, 0, 0, pd - r)) if pd != r else seg)
            off += r
        assert off == w.shape[1], (off, w.shape)
        w = parts[0] if len(parts) == 1 else torch.cat(parts, 1)
        extra = cout_pad - w.shape[0]
        self.weight = F.pad(w, (0, 0, 0, 0, 0, 0, 0, extra)) if extra else w
        self.bias = F.pad(b, (0, extra)) if extra else b
        self.in_channels, self.out_channels = self.weight.shape[1], cout_pad
        self.kernel_size = conv.kernel_size


def effective_convs(convs, x, src1, first_input, seg_real=None):
    """The conv containers the kernels see: `convs` themselves when every channel count is a multiple of 64 (all
    reference scripts), ConvView proxies with zero-padded parameters otherwise."""
    if all(c.out_channels == pad64(c.out_channels) for c in convs) \
            and (first_input is not None or convs[0].in_channels % 64 == 0) \
            and (x is None or x.shape[3] == (convs[0].in_channels - (0 if src1 is None else src1.shape[3]))):
        return convs
    out, prev_real, prev_pad = [], None, None
    for j, c in enumerate(convs):
        cout_pad = pad64(c.out_channels)
        if j == 0 and first_input is not None:
            segs_real = segs_pad = (c.in_channels,)
        elif j == 0:
            segs_pad = (x.shape[3],) + (() if src1 is None else (src1.shape[3],))
            if seg_real is None:
                seg_real = (c.in_channels,) if src1 is None else None
            if seg_real is None:
                raise RuntimeError("channel-padded concat needs the real channel counts of its two segments")
            segs_real = tuple(seg_real)
        else:
            segs_real, segs_pad = (prev_real,), (prev_pad,)
        out.append(ConvView(c, segs_real, segs_pad, cout_pad))
        prev_real, prev_pad = c.out_channels, cout_pad
    return out


def _check_spatial(h, w, levels):
    div = 2 ** (levels - 1)
    if w % div:
        # same exception type as the reference's width check (unet_blocks.py:55)
        raise AssertionError(f"width {w} is not divisible by {div}: up/bridge widths would differ")
    if h % div:
        raise RuntimeError(f"height {h} is not divisible by {div}: Sizes of tensors must match except in dimension 1")


def run_conv_stack(convs, x, first_input=None, pool_last=False, keep_full=True, src1=None, seg_real=None):
    """Runs conv3x3+ReLU for every nn.Conv2d container in `convs`.

    x: NHWC 16-bit input of the first conv, or None when `first_input` = (x0, x1) fp32 planes feed a
    cin<=2 first layer.  src1: optional second K-segment (channel concat) of the first conv; seg_real: the real
    (unpadded) channel counts of (x, src1) when the net's widths are not multiples of 64.
    Returns (full, pooled) of the last conv (channel counts padded to multiples of 64).
    """
    from ..autograd_ops import _needs_grad
    convs = effective_convs(convs, x, src1, first_input, seg_real)
    plist = [t for c in convs for t in (c.weight, c.bias)]
    if _needs_grad(x, src1, *plist):
        # training: the whole block is one autograd node (cross-layer fusion in its backward)
        from ..training import conv_stack_train
        return conv_stack_train(convs, x, src1, first_input, pool_last)
    full, pooled = x, None
    for j, conv in enumerate(convs):
        last = j == len(convs) - 1
        if j == 0 and first_input is not None:
            full = conv3x3_first_op(first_input[0], first_input[1], conv.weight, conv.bias)
            if last and pool_last:
                pooled = avgpool2_op(full)
            continue
        full, pooled = conv3x3_op(full, src1 if j == 0 else None, conv, relu=True,
                                  want_full=(keep_full or not last), want_pool=(last and pool_last))
    return full, pooled


class DownConvBlock(nn.Module):
    """(optional 2x2 average pool ->) 3 x [conv3x3 + ReLU]; reference: unet_blocks.py:7-31."""

    def __init__(self, input_dim, output_dim, initializers, padding, pool=True):
        super().__init__()
        if not padding:
            raise NotImplementedError("only padding=True (the reference's sole configuration) is implemented")
        mods = []
        if pool:
            mods.append(nn.AvgPool2d(kernel_size=2, stride=2, padding=0, ceil_mode=True))
        dims = [input_dim, output_dim, output_dim, output_dim]
        for j in range(3):
            mods.append(nn.Conv2d(dims[j], dims[j + 1], kernel_size=3, stride=1, padding=1))
            mods.append(nn.ReLU(inplace=True))
        self.layers = nn.Sequential(*mods)
        self.layers.apply(init_weights)
        self.pool = pool
        self.input_dim = input_dim

    def convs(self):
        return [m for m in self.layers if isinstance(m, nn.Conv2d)]

    def forward(self, patch):
        """patch: logical NCHW.  Stand-alone use; Unet.forward fuses the pool into the producer instead."""
        if self.input_dim <= 2:
            planes = patch.float()
            x1 = planes[:, 1:2].contiguous() if self.input_dim == 2 else None
            if self.pool:
                raise NotImplementedError("pooled block with <=2 input channels does not occur in the reference")
            full, _ = run_conv_stack(self.convs(), None, first_input=(planes[:, 0:1].contiguous(), x1))
            return as_nchw(full)[:, :self.convs()[-1].out_channels]
        x = pad_channels(as_nhwc(patch))
        if self.pool:
            x = avgpool2_op(x)
        full, _ = run_conv_stack(self.convs(), x, seg_real=(patch.shape[1],))
        return as_nchw(full)[:, :self.convs()[-1].out_channels]


class UpConvBlock(nn.Module):
    """bilinear x2 (align_corners=True) + channel concat [up, bridge] + 3 x [conv3x3 + ReLU];
    reference: unet_blocks.py:34-59.  The concat is never materialised: the first conv reads the
    upsampled tensor and the bridge as two K segments."""

    def __init__(self, input_dim, output_dim, initializers, padding, bilinear=True):
        super().__init__()
        if not bilinear:
            raise NotImplementedError("ConvTranspose2d up-path is dead code in the reference (bilinear=True always)")
        self.bilinear = bilinear
        self.conv_block = DownConvBlock(input_dim, output_dim, initializers, padding, pool=False)

    def forward_nhwc(self, x, bridge, seg_real=None):
        from ..autograd_ops import _needs_grad, packed_weight
        convs = self.conv_block.convs()
        plain = x.shape[3] + bridge.shape[3] == convs[0].in_channels and \
            all(c.out_channels == pad64(c.out_channels) for c in convs)  # (no zero-padded widths in this block)
        if plain and not _needs_grad(x, bridge, *[t for c in convs for t in (c.weight, c.bias)]) \
                and ops.can_fuse_upsample(x, bridge):
            # no-grad path: the first conv interpolates its first K segment itself (bit-identical to the two-kernel
            # form below); under autograd the up-sampled tensor is needed by the weight gradient anyway
            first = ops.conv3x3_up(x, bridge, packed_weight(convs[0], dtype=bridge.dtype), convs[0].bias.detach())
            full, _ = run_conv_stack(convs[1:], first)
            return full
        up = upsample2x_op(x)
        assert up.shape[2] == bridge.shape[2]  # widths (unet_blocks.py:55)
        if up.shape[1] != bridge.shape[1]:
            raise RuntimeError("Sizes of tensors must match except in dimension 1")
        full, _ = run_conv_stack(self.conv_block.convs(), up, src1=bridge, seg_real=seg_real)
        return full

    def forward(self, x, bridge):
        seg_real = (x.shape[1], bridge.shape[1])
        out = self.forward_nhwc(pad_channels(as_nhwc(x)), pad_channels(as_nhwc(bridge)), seg_real=seg_real)
        return as_nchw(out)[:, :self.conv_block.convs()[-1].out_channels]
