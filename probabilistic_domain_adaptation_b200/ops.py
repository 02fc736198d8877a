"""Torch-tensor wrappers over the C ABI.  torch is used for device memory and streams only.

Internal activation layout: contiguous (B, H, W, C) bf16 tensors ("NHWC").  Images, labels and
single-channel outputs are fp32 (B, 1, H, W) == (B, H, W, 1).
"""
import os
import warnings

import torch

from . import _lib

# Storage format of the NHWC activations on the NO-GRAD path (Monte-Carlo inference, teacher, prediction): fp16 keeps
# 11 mantissa bits per layer instead of bf16's 8, which is what brings the sampled logits of the 21-layer trunk within
# the 1e-2 tolerance at full image sizes (tests/test_gpu_baseline_shapes.py); it is also the dtype the reference's own
# student forward runs in under torch_em's autocast.  The training path always uses bf16 (its activation gradients
# need the fp32 exponent range).  PDA_INFER_DTYPE=bf16 selects bf16 everywhere; the fp16 range guard below switches to
# bf16 by itself if an activation ever exceeds +-65504.
INFER_DTYPE = torch.bfloat16 if os.environ.get("PDA_INFER_DTYPE", "fp16").lower() in ("bf16", "bfloat16") \
    else torch.float16
TRAIN_DTYPE = torch.bfloat16
_ACT_DTYPES = (torch.bfloat16, torch.float16)


# When set to a list, every tensor-core conv / fused-Fcomb launch appends (kind, start_event, end_event, work)
# so that bench.py can time the dominant kernel live, on the launching stream, inside the timed region.
PROFILE = None

# Debug switch for the parity tests only: route every conv3x3 through the plain CUDA-core cross-check kernel
# (same bf16 operands, fp32 accumulate) to separate tensor-core kernel bugs from bf16 rounding effects.
FORCE_SIMT_CONV = False


def _stream():
    # raw cudaStream_t of torch's current stream (torch.cuda.current_stream() builds a Stream object: ~10x slower, and
    # this runs once per launch)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


class _Timed:
    def __init__(self, kind, work):
        self.kind, self.work = kind, work

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.kind, self.e0, self.e1, self.work))
        return False


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.PdaError("libpda_b200 kernels need CUDA tensors; this package has no CPU fallback")


# ------------------------------------------------------------------------------------------------
# fp16 range guard: one sticky device flag per GPU, set by the conv epilogues when an fp16 store saturated
# ------------------------------------------------------------------------------------------------
_RANGE = {}


def _range_state(dev):
    st = _RANGE.get(dev)
    if st is None:
        st = {"flag": torch.zeros(1, dtype=torch.int32, device=dev), "host": torch.zeros(1, dtype=torch.int32).pin_memory(),
              "event": torch.cuda.Event(), "pending": False}
        _RANGE[dev] = st
    return st


def range_flag(dev):
    """The device int32[1] the fp16 kernels OR a 1 into when a value left the fp16 range (sticky)."""
    return _range_state(dev)["flag"]


def _range_exceeded(dev):
    global INFER_DTYPE
    _range_state(dev)["flag"].zero_()
    if INFER_DTYPE == torch.float16:
        INFER_DTYPE = torch.bfloat16
        warnings.warn("libpda_b200: an activation exceeded the fp16 range (+-65504) on the no-grad path; the stored value "
                      "was clipped.  Switching the no-grad path to bf16 activations for the rest of this process "
                      "(PDA_INFER_DTYPE=bf16 selects that from the start).", RuntimeWarning)


def check_fp16_range(dev=None):
    """Synchronous check of the sticky flag(s); returns True when every fp16 activation so far was in range.  On a
    violation the no-grad path switches to bf16 (with a RuntimeWarning) and the flag is cleared."""
    ok = True
    for d in ([dev] if dev is not None else list(_RANGE)):
        if int(_range_state(d)["flag"].item()) != 0:
            _range_exceeded(d)
            ok = False
    return ok


def poll_fp16_range(dev):
    """Asynchronous form, called at the start of every no-grad forward: looks at the copy of the flag that an earlier call
    put in flight (no host synchronisation) and enqueues the next copy.  A violation is therefore noticed one or two
    forwards later; the values of the offending forward were clipped to +-65504, never inf."""
    if INFER_DTYPE != torch.float16 or torch.cuda.is_current_stream_capturing():
        return
    st = _range_state(dev)
    if st["pending"] and st["event"].query():
        st["pending"] = False
        if int(st["host"][0]) != 0:
            _range_exceeded(dev)
    if not st["pending"]:
        st["host"].copy_(st["flag"], non_blocking=True)
        st["event"].record()
        st["pending"] = True


def _is_f16(t):
    if t.dtype not in _ACT_DTYPES:
        raise _lib.PdaError(f"NHWC activations must be bf16 or fp16, got {t.dtype}")
    return int(t.dtype == torch.float16)


def pack_conv3x3_weights(w, rot180=False, dtype=torch.bfloat16):
    """(cout, cin, 3, 3) fp32 -> (cout, 9*cin) bf16 / fp16 K-major  [rot180: (cin, 9*cout)]."""
    _need_cuda(w)
    lib = _lib.load()
    cout, cin = w.shape[0], w.shape[1]
    w = w.detach().contiguous().float()
    out = torch.empty((cin, 9 * cout) if rot180 else (cout, 9 * cin), dtype=dtype, device=w.device)
    _lib.check(lib.pda_pack_conv3x3_weights(w.data_ptr(), out.data_ptr(), cout, cin, int(rot180),
                                            int(dtype == torch.float16), _stream()), "pack_conv3x3_weights")
    return out


def conv3x3_first(x0, x1, w, bias, relu=True, dtype=None):
    """x0 (and optional x1): fp32 (B,1,H,W); w: (cout, cin, 3, 3) fp32 -> (B,H,W,cout) NHWC activations.
    dtype: storage format of the output (default: INFER_DTYPE, the no-grad format; the training path passes bf16)."""
    _need_cuda(x0, x1, w, bias)
    lib = _lib.load()
    B, _, H, W = x0.shape
    cout = w.shape[0]
    x0 = x0.contiguous().float()
    x1 = None if x1 is None else x1.contiguous().float()
    dtype = INFER_DTYPE if dtype is None else dtype
    out = torch.empty((B, H, W, cout), dtype=dtype, device=x0.device)
    _lib.check(lib.pda_conv3x3_first(x0.data_ptr(), _ptr(x1), w.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                     B, H, W, cout, int(relu), int(dtype == torch.float16), _stream()),
               "conv3x3_first")
    return out


def conv3x3(src0, src1, w_packed, bias, relu=True, want_full=True, want_pool=False, bn_tile=0, simt=False,
            relu_mask=None):
    """src0/src1: NHWC bf16 or fp16 (src1 optional, channel-concatenated after src0); returns (full, pooled) in the
    same format; w_packed must be packed in that format too.
    relu_mask (B,H,W,cout) bf16: the output is zeroed where relu_mask <= 0 (fused ReLU backward, dgrad use)."""
    _need_cuda(src0, src1, w_packed, bias)
    lib = _lib.load()
    B, H, W, c0 = src0.shape
    c1 = 0 if src1 is None else src1.shape[3]
    cout = w_packed.shape[0]
    f16 = _is_f16(src0)
    assert w_packed.shape[1] == 9 * (c0 + c1), (w_packed.shape, c0, c1)
    assert src0.is_contiguous() and (src1 is None or src1.is_contiguous())
    assert w_packed.dtype == src0.dtype and (src1 is None or src1.dtype == src0.dtype), "mixed activation formats"
    full = torch.empty((B, H, W, cout), dtype=src0.dtype, device=src0.device) if want_full else None
    pool = torch.empty((B, H // 2, W // 2, cout), dtype=src0.dtype, device=src0.device) if want_pool else None
    if relu_mask is not None:
        assert relu_mask.shape == (B, H, W, cout) and relu_mask.is_contiguous() and relu_mask.dtype == torch.bfloat16
    if (simt or FORCE_SIMT_CONV) and f16:
        raise _lib.PdaError("the CUDA-core cross-check conv is bf16 only (set PDA_INFER_DTYPE=bf16)")
    if simt or FORCE_SIMT_CONV:
        rc = lib.pda_conv3x3_bf16_simt(src0.data_ptr(), c0, _ptr(src1), c1, w_packed.data_ptr(), _ptr(bias),
                                       _ptr(full), _ptr(pool), B, H, W, cout, int(relu), _stream())
        if rc == 0 and relu_mask is not None:  # the cross-check conv has no fused mask epilogue
            full = relu_pool_bwd(full, None, relu_mask)
    else:
        with _Timed("conv3x3_tc", 2.0 * 9 * (c0 + c1) * cout * B * H * W):
            rc = lib.pda_conv3x3_tc(src0.data_ptr(), c0, _ptr(src1), c1, w_packed.data_ptr(), _ptr(bias),
                                    _ptr(full), _ptr(pool), _ptr(relu_mask), B, H, W, cout, int(relu),
                                    int(bn_tile), f16, range_flag(src0.device).data_ptr() if f16 else 0, _stream())
    _lib.check(rc, "conv3x3")
    return full, pool


# The first conv of an up-path block consumes the bilinear x2 of its low-resolution input directly (no-grad path):
# PDA_FUSE_UPSAMPLE=0 keeps the separate up-sampling kernel (the results are bit-identical).
FUSE_UPSAMPLE = os.environ.get("PDA_FUSE_UPSAMPLE", "1") != "0"


_FUSE_DEBUG_MASK = int(os.environ["PDA_FUSE_MASK"]) if "PDA_FUSE_MASK" in os.environ else None


def can_fuse_upsample(x_low, bridge):
    """The fused form lives in the CTA-pair conv kernel (needs >= 2 pixel tiles of 8 x 16/32 px)."""
    if not FUSE_UPSAMPLE or _lib.load().pda_set_conv_pair(-1) != 1:
        return False
    B, H, W, c1 = bridge.shape
    if _FUSE_DEBUG_MASK is not None and not (_FUSE_DEBUG_MASK & (c1 // 64)):   # PDA_FUSE_MASK: debugging aid that fuses
        return False                                                           # only the layers with 64 * mask bridge channels
    return x_low.shape[1] * 2 == H and x_low.shape[2] * 2 == W and B * ((W + 7) // 8) * ((H + 31) // 32) >= 2


def conv3x3_up(x_low, bridge, w_packed, bias, relu=True):
    """conv3x3 over cat(bilinear_x2(x_low), bridge) without materialising the up-sampled tensor."""
    _need_cuda(x_low, bridge, w_packed, bias)
    lib = _lib.load()
    B, H, W, c1 = bridge.shape
    c0, cout = x_low.shape[3], w_packed.shape[0]
    f16 = _is_f16(bridge)
    assert x_low.dtype == bridge.dtype == w_packed.dtype and x_low.is_contiguous() and bridge.is_contiguous()
    assert x_low.shape[:3] == (B, H // 2, W // 2) and w_packed.shape[1] == 9 * (c0 + c1)
    full = torch.empty((B, H, W, cout), dtype=bridge.dtype, device=bridge.device)
    with _Timed("conv3x3_tc", 2.0 * 9 * (c0 + c1) * cout * B * H * W):
        rc = lib.pda_conv3x3_up_tc(x_low.data_ptr(), c0, bridge.data_ptr(), c1, w_packed.data_ptr(), _ptr(bias),
                                   full.data_ptr(), 0, B, H, W, cout, int(relu), f16,
                                   range_flag(bridge.device).data_ptr() if f16 else 0, _stream())
    _lib.check(rc, "conv3x3_up")
    return full


def avgpool2(x):
    _need_cuda(x)
    lib = _lib.load()
    B, H, W, C = x.shape
    out = torch.empty((B, H // 2, W // 2, C), dtype=x.dtype, device=x.device)
    _lib.check(lib.pda_avgpool2(x.data_ptr(), out.data_ptr(), B, H, W, C, _is_f16(x), _stream()), "avgpool2")
    return out


def upsample2x(x):
    _need_cuda(x)
    lib = _lib.load()
    B, h, w, C = x.shape
    out = torch.empty((B, 2 * h, 2 * w, C), dtype=x.dtype, device=x.device)
    _lib.check(lib.pda_upsample2x_bilinear(x.data_ptr(), out.data_ptr(), B, h, w, C, _is_f16(x), _stream()),
               "upsample2x")
    return out


def gauss_head(enc, w_head, b_head, latent):
    """enc: (B,h,w,C) bf16; w_head (2L, C, 1, 1) fp32 -> (B, 2L) fp32 = (mu | log_sigma)."""
    _need_cuda(enc, w_head, b_head)
    lib = _lib.load()
    B, h, w, C = enc.shape
    P = h * w
    rows = lib.pda_gauss_head_scratch_rows(P)
    scratch = torch.empty((B, rows, C), dtype=torch.float32, device=enc.device)
    out = torch.empty((B, 2 * latent), dtype=torch.float32, device=enc.device)
    _lib.check(lib.pda_gauss_head(enc.data_ptr(), w_head.data_ptr(), b_head.data_ptr(), scratch.data_ptr(),
                                  out.data_ptr(), B, P, C, latent, _is_f16(enc), _stream()), "gauss_head")
    return out


def latent_samples(mu_logsigma, eps):
    """eps (S,B,L) -> z (S,B,L) = mu + exp(log_sigma) * eps."""
    _need_cuda(mu_logsigma, eps)
    lib = _lib.load()
    S, B, L = eps.shape
    eps = eps.contiguous().float()
    z = torch.empty_like(eps)
    _lib.check(lib.pda_latent_samples(mu_logsigma.data_ptr(), eps.data_ptr(), z.data_ptr(), S, B, L, _stream()),
               "latent_samples")
    return z


def kl_diag_gauss(mls_q, mls_p):
    _need_cuda(mls_q, mls_p)
    lib = _lib.load()
    B, L2 = mls_q.shape
    kl = torch.empty((B,), dtype=torch.float32, device=mls_q.device)
    _lib.check(lib.pda_kl_diag_gauss(mls_q.data_ptr(), mls_p.data_ptr(), kl.data_ptr(), B, L2 // 2, _stream()),
               "kl_diag_gauss")
    return kl


def fcomb_mc_consensus(feat, z, w1, b1, w2, b2, w3, b3, upper=0.9, lower=0.1, want_mean=True, want_weight=True,
                       want_mask=False, want_logits=False, want_probs=False, precision="bf16"):
    """feat (B,H,W,64) bf16 or fp16; z (S,B,L) fp32.  Returns dict of (B,1,H,W) / (S,B,1,H,W) tensors.
    "range_flag" (precision "bf16" only): the call's scratch; element 0 is the int32 fp16-range flag that the kernel
    raises (and acts on, by re-running the batch in fp32 on the device) -- `fcomb_bwd` takes it as `fwd_flag`."""
    _need_cuda(feat, z, w1)
    lib = _lib.load()
    B, H, W, C = feat.shape
    S, Bz, L = z.shape
    assert Bz == B and C == 64 and w1.shape[1] == C + L and feat.is_contiguous()
    dev = feat.device
    P = H * W
    mean = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev) if want_mean else None
    weight = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev) if want_weight else None
    mask = torch.empty((B, 1, H, W), dtype=torch.int64, device=dev) if want_mask else None
    logits = torch.empty((S, B, 1, H, W), dtype=torch.float32, device=dev) if want_logits else None
    probs = torch.empty((S, B, 1, H, W), dtype=torch.float32, device=dev) if want_probs else None
    z = z.contiguous().float()
    scratch = None
    args = (feat.data_ptr(), z.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
            w3.data_ptr(), b3.data_ptr(), B, P, S, L, float(upper), float(lower), _ptr(mean), _ptr(weight),
            _ptr(mask), _ptr(logits), _ptr(probs))
    with _Timed("fcomb_mc", float(B * P)):
        if precision == "bf16":
            # per-call scratch from torch's caching allocator: nothing is shared between launches, streams or graphs
            scratch = torch.empty(lib.pda_fcomb_scratch_floats(S, B), dtype=torch.float32, device=dev)
            rc = lib.pda_fcomb_mc_consensus(*args, scratch.data_ptr(), _is_f16(feat), _stream())
        else:
            rc = lib.pda_fcomb_mc_consensus_fp32(*args, _is_f16(feat), _stream())
    _lib.check(rc, "fcomb_mc_consensus")
    return {"mean": mean, "weight": weight, "mask": mask, "logits": logits, "probs": probs, "range_flag": scratch}


def fcomb_mc_consensus_deep(feat, z, w1, b1, wmid, bmid, w3, b3, upper=0.9, lower=0.1, want_mean=True, want_weight=True,
                            want_mask=False, want_logits=False, want_probs=False):
    """General-depth Fcomb (wmid (n_mid, 64, 64) or None): plain fp32 kernel, same outputs as fcomb_mc_consensus."""
    _need_cuda(feat, z, w1)
    lib = _lib.load()
    B, H, W, C = feat.shape
    S, Bz, L = z.shape
    assert Bz == B and C == 64 and w1.shape[1] == C + L and feat.is_contiguous()
    dev = feat.device
    P = H * W
    mean = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev) if want_mean else None
    weight = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev) if want_weight else None
    mask = torch.empty((B, 1, H, W), dtype=torch.int64, device=dev) if want_mask else None
    logits = torch.empty((S, B, 1, H, W), dtype=torch.float32, device=dev) if want_logits else None
    probs = torch.empty((S, B, 1, H, W), dtype=torch.float32, device=dev) if want_probs else None
    z = z.contiguous().float()
    n_mid = 0 if wmid is None else wmid.shape[0]
    w1, b1, w3, b3 = (t.contiguous().float() for t in (w1, b1, w3, b3))
    _lib.check(lib.pda_fcomb_mc_consensus_deep(feat.data_ptr(), z.data_ptr(), w1.data_ptr(), b1.data_ptr(), _ptr(wmid),
                                               _ptr(bmid), n_mid, w3.data_ptr(), b3.data_ptr(), B, P, S, L, float(upper),
                                               float(lower), _ptr(mean), _ptr(weight), _ptr(mask), _ptr(logits),
                                               _ptr(probs), _is_f16(feat), _stream()), "fcomb_mc_consensus_deep")
    return {"mean": mean, "weight": weight, "mask": mask, "logits": logits, "probs": probs, "range_flag": None}


_EMA_CHUNK = 16384  # elements per 256-thread block of the multi-tensor kernels (<= 65536)


def build_ema_table(teacher_params, student_params):
    """Device int64 table (n_chunks, 3) = (teacher_ptr, student_ptr, numel) for pda_multi_tensor_ema."""
    rows = []
    for t, s in zip(teacher_params, student_params):
        assert t.dtype == torch.float32 and s.dtype == torch.float32 and t.numel() == s.numel()
        assert t.is_contiguous() and s.is_contiguous()
        n = t.numel()
        for off in range(0, n, _EMA_CHUNK):
            rows.append((t.data_ptr() + 4 * off, s.data_ptr() + 4 * off, min(_EMA_CHUNK, n - off)))
    dev = teacher_params[0].device
    return torch.tensor(rows, dtype=torch.int64).to(dev)


def multi_tensor_ema(table, momentum, iteration_dev=None):
    """iteration_dev (int64 device scalar): AdaMT warm-up momentum min(1 - 1/(it+1), momentum) evaluated on the device;
    the counter is incremented by the call."""
    _need_cuda(table, iteration_dev)
    lib = _lib.load()
    if iteration_dev is not None:
        assert iteration_dev.dtype == torch.int64 and iteration_dev.numel() == 1
        _lib.check(lib.pda_multi_tensor_ema_warmup(table.data_ptr(), table.shape[0], float(momentum),
                                                   iteration_dev.data_ptr(), _stream()), "multi_tensor_ema_warmup")
        return
    _lib.check(lib.pda_multi_tensor_ema(table.data_ptr(), table.shape[0], float(momentum), _stream()),
               "multi_tensor_ema")


# ------------------------------------------------------------------------------------------------
# training (backward) wrappers
# ------------------------------------------------------------------------------------------------
# Self-cleaning per-layer scratch for the weight-gradient kernel (no memset per launch).  Off: measured slower than a
# fresh scratch + memset (the clean-up stores double the re-layout kernel's write traffic).
WGRAD_PERSISTENT_SCRATCH = False


# PDA_WGRAD_DETERMINISTIC=1 (or ops.WGRAD_DETERMINISTIC = True): the weight-gradient kernel stores per-CTA partial sums and a
# second kernel adds them in a fixed order instead of fp32 atomics -- bit-identical gradients from run to run (the
# conv weight gradients are the only non-deterministic sums of a training step besides the Fcomb backward's)
WGRAD_DETERMINISTIC = os.environ.get("PDA_WGRAD_DETERMINISTIC", "0") == "1"


def wgrad_scratch(owner, ctot, cout, dev):
    """The persistent, self-cleaning scratch of one conv layer's weight-gradient kernel (zeroed once, here; the kernel
    leaves it all zero again after every launch, so no memset runs per step).  Kept on the nn.Conv2d container."""
    n = _lib.load().pda_conv3x3_wgrad_scratch_floats(ctot, cout)
    cached = owner.__dict__.get("_pda_wgrad_scratch")
    if cached is None or cached.numel() != n or cached.device != dev:
        cached = owner.__dict__["_pda_wgrad_scratch"] = torch.zeros(n, dtype=torch.float32, device=dev)
    return cached


def conv3x3_wgrad(src0, src1, dz, want_bias=True, owner=None):
    """dW (cout, c0+c1, 3, 3) fp32 and db (cout,) of conv3x3 from its NHWC bf16 inputs and dZ.
    owner: the nn.Conv2d container of the layer -- its persistent zero-invariant scratch is used (no memset per call)."""
    _need_cuda(src0, src1, dz)
    lib = _lib.load()
    B, H, W, c0 = src0.shape
    c1 = 0 if src1 is None else src1.shape[3]
    cout = dz.shape[3]
    assert dz.shape[:3] == src0.shape[:3] and dz.is_contiguous() and src0.is_contiguous()
    assert src0.dtype == torch.bfloat16 and dz.dtype == torch.bfloat16
    dev = src0.device
    if WGRAD_DETERMINISTIC:
        scratch = torch.empty(lib.pda_conv3x3_wgrad_det_scratch_floats(c0 + c1, cout, B, H, W), dtype=torch.float32,
                              device=dev)
        dw = torch.empty((cout, c0 + c1, 3, 3), dtype=torch.float32, device=dev)
        db = torch.empty((cout,), dtype=torch.float32, device=dev) if want_bias else None
        with _Timed("wgrad3x3_tc", 2.0 * 9 * (c0 + c1) * cout * B * H * W):
            rc = lib.pda_conv3x3_wgrad_bf16_det(src0.data_ptr(), c0, _ptr(src1), c1, dz.data_ptr(), scratch.data_ptr(),
                                                dw.data_ptr(), _ptr(db), B, H, W, cout, 0, _stream())
        _lib.check(rc, "conv3x3_wgrad_det")
        return dw, db
    if owner is not None and WGRAD_PERSISTENT_SCRATCH:
        scratch, is_zero = wgrad_scratch(owner, c0 + c1, cout, dev), 1
    else:
        scratch = torch.empty(lib.pda_conv3x3_wgrad_scratch_floats(c0 + c1, cout), dtype=torch.float32, device=dev)
        is_zero = 0
    dw = torch.empty((cout, c0 + c1, 3, 3), dtype=torch.float32, device=dev)
    db = torch.empty((cout,), dtype=torch.float32, device=dev) if want_bias else None
    with _Timed("wgrad3x3_tc", 2.0 * 9 * (c0 + c1) * cout * B * H * W):
        rc = lib.pda_conv3x3_wgrad_bf16(src0.data_ptr(), c0, _ptr(src1), c1, dz.data_ptr(), scratch.data_ptr(),
                                        dw.data_ptr(), _ptr(db), B, H, W, cout, 0, is_zero, _stream())
    _lib.check(rc, "conv3x3_wgrad")
    return dw, db


def relu_pool_bwd(dfull, dpool, y, shape=None, want_bias=False):
    """dZ = (dFull + 0.25 * up2(dPool)) * (y > 0); y None = no ReLU mask.  NHWC bf16.
    want_bias: also returns db = sum over pixels of dZ (the conv bias gradient) -> (dz, db)."""
    _need_cuda(dfull, dpool, y)
    lib = _lib.load()
    ref = y if y is not None else dfull
    B, H, W, C = ref.shape if ref is not None else shape
    dev = (ref if ref is not None else dpool).device
    dz = torch.empty((B, H, W, C), dtype=torch.bfloat16, device=dev)
    db = torch.empty((C,), dtype=torch.float32, device=dev) if want_bias else None
    _lib.check(lib.pda_relu_pool_bwd_bf16(_ptr(dfull), _ptr(dpool), _ptr(y), dz.data_ptr(), _ptr(db), B, H, W, C,
                                          _stream()), "relu_pool_bwd")
    return (dz, db) if want_bias else dz


def upsample2x_bwd(dout):
    _need_cuda(dout)
    lib = _lib.load()
    B, H2, W2, C = dout.shape
    din = torch.empty((B, H2 // 2, W2 // 2, C), dtype=torch.bfloat16, device=dout.device)
    _lib.check(lib.pda_upsample2x_bilinear_bwd_bf16(dout.data_ptr(), din.data_ptr(), B, H2 // 2, W2 // 2, C,
                                                    _stream()), "upsample2x_bwd")
    return din


def conv3x3_first_bwd(x0, x1, out, dout, premasked=False):
    """premasked: dout already carries the layer's ReLU mask (output of the next layer's dgrad conv with relu_mask=out):
    the forward output is then not read at all."""
    _need_cuda(x0, x1, out, dout)
    lib = _lib.load()
    B, H, W, cout = dout.shape
    cin = 1 if x1 is None else 2
    dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=dout.device)
    db = torch.empty((cout,), dtype=torch.float32, device=dout.device)
    scratch = torch.empty(lib.pda_conv3x3_first_bwd_scratch_floats(B, H, W, cout, cin), dtype=torch.float32,
                          device=dout.device)
    _lib.check(lib.pda_conv3x3_first_bwd(x0.data_ptr(), _ptr(x1), 0 if premasked else out.data_ptr(), dout.data_ptr(),
                                         dw.data_ptr(),
                                         db.data_ptr(), B, H, W, cout, scratch.data_ptr(), _stream()),
               "conv3x3_first_bwd")
    return dw, db


def gauss_head_fwd_train(enc, w_head, b_head, latent):
    """Like gauss_head but also returns the stage-1 partial sums needed by the backward."""
    _need_cuda(enc, w_head, b_head)
    lib = _lib.load()
    B, h, w, C = enc.shape
    P = h * w
    rows = lib.pda_gauss_head_scratch_rows(P)
    scratch = torch.empty((B, rows, C), dtype=torch.float32, device=enc.device)
    out = torch.empty((B, 2 * latent), dtype=torch.float32, device=enc.device)
    _lib.check(lib.pda_gauss_head(enc.data_ptr(), w_head.data_ptr(), b_head.data_ptr(), scratch.data_ptr(),
                                  out.data_ptr(), B, P, C, latent, _is_f16(enc), _stream()), "gauss_head")
    return out, scratch


def gauss_head_bwd(dmls, w_head, scratch, enc, latent):
    _need_cuda(dmls, w_head, scratch, enc)
    lib = _lib.load()
    B, h, w, C = enc.shape
    P = h * w
    dev = enc.device
    mean = torch.empty((B, C), dtype=torch.float32, device=dev)
    _lib.check(lib.pda_gauss_head_mean(scratch.data_ptr(), mean.data_ptr(), B, P, C, _stream()), "gauss_head_mean")
    dw = torch.empty((2 * latent, C, 1, 1), dtype=torch.float32, device=dev)
    db = torch.empty((2 * latent,), dtype=torch.float32, device=dev)
    dmean = torch.empty((B, C), dtype=torch.float32, device=dev)
    denc = torch.empty_like(enc)
    dmls = dmls.contiguous().float()
    _lib.check(lib.pda_gauss_head_bwd(dmls.data_ptr(), w_head.data_ptr(), mean.data_ptr(), enc.data_ptr(),
                                      dw.data_ptr(), db.data_ptr(), dmean.data_ptr(), denc.data_ptr(), B, P, C,
                                      latent, _stream()), "gauss_head_bwd")
    return denc, dw, db


def kl_diag_gauss_bwd(mls_q, mls_p, dkl):
    _need_cuda(mls_q, mls_p, dkl)
    lib = _lib.load()
    B, L2 = mls_q.shape
    dq, dp = torch.empty_like(mls_q), torch.empty_like(mls_p)
    dkl = dkl.contiguous().float()
    _lib.check(lib.pda_kl_diag_gauss_bwd(mls_q.data_ptr(), mls_p.data_ptr(), dkl.data_ptr(), dq.data_ptr(),
                                         dp.data_ptr(), B, L2 // 2, _stream()), "kl_diag_gauss_bwd")
    return dq, dp


def _consm_ptrs(consm):
    if consm is None:
        return 0, 0
    if consm.dtype == torch.int64:
        return 0, consm.data_ptr()
    assert consm.dtype == torch.float32
    return consm.data_ptr(), 0


def recon_loss_fwd(logits, segm, consm, dice):
    """-> out2 (sum, mean) fp32, stats3 fp32 (kept for backward)."""
    _need_cuda(logits, segm, consm)
    lib = _lib.load()
    n = logits.numel()
    assert segm.numel() == n and (consm is None or consm.numel() == n)
    dev = logits.device
    partial = torch.empty(3 * lib.pda_recon_loss_blocks(n), dtype=torch.float64, device=dev)
    out2 = torch.empty(2, dtype=torch.float32, device=dev)
    stats = torch.empty(3, dtype=torch.float32, device=dev)
    cf, ci = _consm_ptrs(consm)
    _lib.check(lib.pda_recon_loss_fwd(logits.data_ptr(), segm.data_ptr(), cf, ci, n, int(dice), partial.data_ptr(),
                                      out2.data_ptr(), stats.data_ptr(), _stream()), "recon_loss_fwd")
    return out2, stats


def recon_loss_bwd(logits, segm, consm, dice, stats, gout2):
    _need_cuda(logits, segm, consm, stats, gout2)
    lib = _lib.load()
    n = logits.numel()
    dlogits = torch.empty_like(logits)
    cf, ci = _consm_ptrs(consm)
    _lib.check(lib.pda_recon_loss_bwd(logits.data_ptr(), segm.data_ptr(), cf, ci, n, int(dice), stats.data_ptr(),
                                      gout2.data_ptr(), dlogits.data_ptr(), _stream()), "recon_loss_bwd")
    return dlogits


def build_l2_tables(params):
    """Device tables for the multi-tensor L2 norm: forward rows (ptr, numel, tensor, 0); backward rows
    (w_ptr, byte offset inside the flat gradient buffer, numel, tensor).  Returns (fwd, bwd, offsets, total)."""
    fwd, bwd, offsets, total = [], [], [], 0
    for t, p in enumerate(params):
        assert p.dtype == torch.float32 and p.is_contiguous()
        n = p.numel()
        offsets.append(total)
        for off in range(0, n, _EMA_CHUNK):
            m = min(_EMA_CHUNK, n - off)
            fwd.append((p.data_ptr() + 4 * off, m, t, 0))
            bwd.append((p.data_ptr() + 4 * off, 4 * (total + off), m, t))
        total += (n + 3) // 4 * 4  # keep every tensor 16-byte aligned inside the flat buffer
    dev = params[0].device
    return (torch.tensor(fwd, dtype=torch.int64).to(dev), torch.tensor(bwd, dtype=torch.int64).to(dev), offsets, total)


def multi_tensor_l2norm_fwd(table, n_tensors):
    lib = _lib.load()
    dev = table.device
    partial = torch.empty(table.shape[0], dtype=torch.float64, device=dev)
    norms = torch.empty(n_tensors, dtype=torch.float32, device=dev)
    out = torch.empty((), dtype=torch.float32, device=dev)
    _lib.check(lib.pda_multi_tensor_l2norm_fwd(table.data_ptr(), table.shape[0], n_tensors, partial.data_ptr(),
                                               norms.data_ptr(), out.data_ptr(), _stream()), "l2norm_fwd")
    return out, norms


def multi_tensor_l2norm_bwd(table, norms, gout, total):
    lib = _lib.load()
    flat = torch.empty(total, dtype=torch.float32, device=table.device)
    gout = gout.contiguous().float()
    _lib.check(lib.pda_multi_tensor_l2norm_bwd(table.data_ptr(), table.shape[0], norms.data_ptr(), gout.data_ptr(),
                                               flat.data_ptr(), _stream()), "l2norm_bwd")
    return flat


def fcomb_bwd(feat, z, w1, b1, w2, b2, w3, dlogit, precision="bf16", fwd_flag=None):
    """Backward of Fcomb for one latent sample: feat (B,H,W,64) bf16, z (B,L), dlogit (B,1,H,W) fp32.
    precision "bf16": tensor-core kernel; "fp32": the exact CUDA-core baseline.  fwd_flag: the "range_flag" tensor of
    the forward call -- when the forward fell back to fp32 (fp16 range guard) the backward does the same, on the device."""
    _need_cuda(feat, z, w1, dlogit)
    lib = _lib.load()
    B, H, W, C = feat.shape
    L = z.shape[1]
    dev = feat.device
    dfeat = torch.empty_like(feat)
    dw1, db1 = torch.empty_like(w1), torch.empty_like(b1)
    dw2, db2 = torch.empty_like(w2), torch.empty_like(b2)
    dw3 = torch.empty_like(w3)
    db3 = torch.empty(1, dtype=torch.float32, device=dev)
    dz = torch.empty((B, L), dtype=torch.float32, device=dev)
    scratch = torch.empty(64 * 64 + 2 * B * 64, dtype=torch.float32, device=dev)
    z = z.contiguous().float()
    dlogit = dlogit.contiguous().float()
    args = (feat.data_ptr(), z.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), w3.data_ptr(),
            dlogit.data_ptr(), B, H * W, L, dfeat.data_ptr(), dw1.data_ptr(), db1.data_ptr(), dw2.data_ptr(),
            db2.data_ptr(), dw3.data_ptr(), db3.data_ptr(), dz.data_ptr(), scratch.data_ptr())
    with _Timed("fcomb_bwd", float(B * H * W)):
        if precision == "bf16":
            rc = lib.pda_fcomb_bwd(*args, _ptr(fwd_flag), _stream())
        else:
            rc = lib.pda_fcomb_bwd_fp32(*args, _stream())
    _lib.check(rc, "fcomb_bwd")
    return dfeat, dw1, db1, dw2, db2, dw3, db3, dz


def dice_score(seg, gt, threshold_seg=None, threshold_gt=None):
    """my_utils/util.py:17-44 on the device: fp32 tensors of equal shape -> 0-dim fp32 tensor (no host sync)."""
    _need_cuda(seg, gt)
    lib = _lib.load()
    assert seg.shape == gt.shape, f"{seg.shape}, {gt.shape}"
    seg, gt = seg.contiguous().float(), gt.contiguous().float()
    n = seg.numel()
    partial = torch.empty(3 * lib.pda_recon_loss_blocks(n), dtype=torch.float64, device=seg.device)
    out = torch.empty(1, dtype=torch.float32, device=seg.device)
    nan = float("nan")
    _lib.check(lib.pda_dice_score(seg.data_ptr(), gt.data_ptr(), n, nan if threshold_seg is None else threshold_seg,
                                  nan if threshold_gt is None else threshold_gt, partial.data_ptr(), out.data_ptr(),
                                  _stream()), "dice_score")
    return out[0]
