"""Torch-tensor wrappers over the C ABI.  torch is used for device memory and streams only.

Internal activation layout: contiguous (B, H, W, C) bf16 tensors ("NHWC").  Images, labels and
single-channel outputs are fp32 (B, 1, H, W) == (B, H, W, 1).
"""
import torch

from . import _lib


# When set to a list, every tensor-core conv / fused-Fcomb launch appends (kind, start_event, end_event, work)
# so that bench.py can time the dominant kernel live, on the launching stream, inside the timed region.
PROFILE = None


def _stream():
    return torch.cuda.current_stream().cuda_stream


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


def pack_conv3x3_weights(w, rot180=False):
    """(cout, cin, 3, 3) fp32 -> (cout, 9*cin) bf16 K-major  [rot180: (cin, 9*cout)]."""
    _need_cuda(w)
    lib = _lib.load()
    cout, cin = w.shape[0], w.shape[1]
    w = w.detach().contiguous().float()
    out = torch.empty((cin, 9 * cout) if rot180 else (cout, 9 * cin), dtype=torch.bfloat16, device=w.device)
    _lib.check(lib.pda_pack_conv3x3_weights(w.data_ptr(), out.data_ptr(), cout, cin, int(rot180), _stream()),
               "pack_conv3x3_weights")
    return out


def conv3x3_first(x0, x1, w, bias, relu=True):
    """x0 (and optional x1): fp32 (B,1,H,W); w: (cout, cin, 3, 3) fp32 -> (B,H,W,cout) bf16."""
    _need_cuda(x0, x1, w, bias)
    lib = _lib.load()
    B, _, H, W = x0.shape
    cout = w.shape[0]
    x0 = x0.contiguous().float()
    x1 = None if x1 is None else x1.contiguous().float()
    out = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device=x0.device)
    _lib.check(lib.pda_conv3x3_first(x0.data_ptr(), _ptr(x1), w.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                     B, H, W, cout, int(relu), _stream()), "conv3x3_first")
    return out


def conv3x3(src0, src1, w_packed, bias, relu=True, want_full=True, want_pool=False, bn_tile=0, simt=False):
    """src0/src1: NHWC bf16 (src1 optional, channel-concatenated after src0); returns (full, pooled)."""
    _need_cuda(src0, src1, w_packed, bias)
    lib = _lib.load()
    B, H, W, c0 = src0.shape
    c1 = 0 if src1 is None else src1.shape[3]
    cout = w_packed.shape[0]
    assert w_packed.shape[1] == 9 * (c0 + c1), (w_packed.shape, c0, c1)
    assert src0.is_contiguous() and (src1 is None or src1.is_contiguous())
    full = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device=src0.device) if want_full else None
    pool = torch.empty((B, H // 2, W // 2, cout), dtype=torch.bfloat16, device=src0.device) if want_pool else None
    if simt:
        rc = lib.pda_conv3x3_bf16_simt(src0.data_ptr(), c0, _ptr(src1), c1, w_packed.data_ptr(), _ptr(bias),
                                       _ptr(full), _ptr(pool), B, H, W, cout, int(relu), _stream())
    else:
        with _Timed("conv3x3_tc", 2.0 * 9 * (c0 + c1) * cout * B * H * W):
            rc = lib.pda_conv3x3_bf16(src0.data_ptr(), c0, _ptr(src1), c1, w_packed.data_ptr(), _ptr(bias),
                                      _ptr(full), _ptr(pool), B, H, W, cout, int(relu), int(bn_tile), _stream())
    _lib.check(rc, "conv3x3")
    return full, pool


def avgpool2(x):
    _need_cuda(x)
    lib = _lib.load()
    B, H, W, C = x.shape
    out = torch.empty((B, H // 2, W // 2, C), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.pda_avgpool2_bf16(x.data_ptr(), out.data_ptr(), B, H, W, C, _stream()), "avgpool2")
    return out


def upsample2x(x):
    _need_cuda(x)
    lib = _lib.load()
    B, h, w, C = x.shape
    out = torch.empty((B, 2 * h, 2 * w, C), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.pda_upsample2x_bilinear_bf16(x.data_ptr(), out.data_ptr(), B, h, w, C, _stream()), "upsample2x")
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
                                  out.data_ptr(), B, P, C, latent, _stream()), "gauss_head")
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
    """feat (B,H,W,64) bf16; z (S,B,L) fp32.  Returns dict of (B,1,H,W) / (S,B,1,H,W) tensors."""
    _need_cuda(feat, z, w1)
    lib = _lib.load()
    B, H, W, C = feat.shape
    S, Bz, L = z.shape
    assert Bz == B and C == 64 and w1.shape[1] == C + L
    dev = feat.device
    P = H * W
    mean = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev) if want_mean else None
    weight = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev) if want_weight else None
    mask = torch.empty((B, 1, H, W), dtype=torch.int64, device=dev) if want_mask else None
    logits = torch.empty((S, B, 1, H, W), dtype=torch.float32, device=dev) if want_logits else None
    probs = torch.empty((S, B, 1, H, W), dtype=torch.float32, device=dev) if want_probs else None
    z = z.contiguous().float()
    fn = lib.pda_fcomb_mc_consensus if precision == "bf16" else lib.pda_fcomb_mc_consensus_fp32
    with _Timed("fcomb_mc", float(B * P)):
        rc = fn(feat.data_ptr(), z.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                w3.data_ptr(), b3.data_ptr(), B, P, S, L, float(upper), float(lower), _ptr(mean), _ptr(weight),
                _ptr(mask), _ptr(logits), _ptr(probs), _stream())
    _lib.check(rc, "fcomb_mc_consensus")
    return {"mean": mean, "weight": weight, "mask": mask, "logits": logits, "probs": probs}


_EMA_CHUNK = 65536


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


def multi_tensor_ema(table, momentum):
    _need_cuda(table)
    lib = _lib.load()
    _lib.check(lib.pda_multi_tensor_ema(table.data_ptr(), table.shape[0], float(momentum), _stream()),
               "multi_tensor_ema")
