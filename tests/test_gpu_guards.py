"""Out-of-bounds guards for the hand-written kernels (compute-sanitizer is closed on this GPU pool:
profiles/r02_sanitizer_closed.md).

Every kernel is called THROUGH THE C ABI with each output carved out of the middle of a larger allocation whose
remainder is filled with a canary pattern; after the launch the canaries must be untouched (out-of-bounds writes) and
the payload must equal what the ordinary op wrappers (own allocations) produce bit for bit (a stale read / partial write
shows up here).  Ragged shapes (sizes that are not multiples of the tile) are used on purpose.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

CANARY = 0x7A  # byte pattern: bf16 / fp16 / fp32 / int64 values made of it are huge and recognisable
PAD = 4096     # bytes of canary on each side


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


class Guarded:
    """A tensor of the requested shape / dtype that lives between two canary regions of one byte buffer."""

    def __init__(self, shape, dtype, dev):
        self.n = int(torch.tensor(shape).prod()) * torch.empty((), dtype=dtype).element_size()
        self.buf = torch.full((self.n + 2 * PAD,), CANARY, dtype=torch.uint8, device=dev)
        self.t = self.buf[PAD:PAD + self.n].view(dtype).view(shape)

    def intact(self):
        return bool((self.buf[:PAD] == CANARY).all()) and bool((self.buf[PAD + self.n:] == CANARY).all())


def _lib():
    from probabilistic_domain_adaptation_b200 import _lib
    return _lib.load(), _lib


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,H,W,c0,c1,cout,pool", [(1, 18, 10, 64, 0, 64, True), (2, 34, 24, 128, 64, 128, False),
                                                   (1, 6, 40, 64, 0, 256, True), (3, 16, 8, 64, 0, 64, False)])
def test_conv3x3_writes_stay_inside_their_tensors(B, H, W, c0, c1, cout, pool, dt, pair):
    from probabilistic_domain_adaptation_b200 import ops
    lib, L = _lib()
    dev = _dev()
    g = torch.Generator().manual_seed(H * W + cout)
    s0 = torch.randn(B, H, W, c0, generator=g).to(dev).to(dt)
    s1 = torch.randn(B, H, W, c1, generator=g).to(dev).to(dt) if c1 else None
    w = (torch.randn(cout, c0 + c1, 3, 3, generator=g) * 0.05).to(dev)
    bias = (torch.randn(cout, generator=g) * 0.1).to(dev)
    wp = ops.pack_conv3x3_weights(w, dtype=dt)
    prev = lib.pda_set_conv_pair(pair)
    try:
        want_full, want_pool = ops.conv3x3(s0, s1, wp, bias, want_full=True, want_pool=pool)
        full = Guarded((B, H, W, cout), dt, dev)
        pooled = Guarded((B, H // 2, W // 2, cout), dt, dev) if pool else None
        f16 = int(dt == torch.float16)
        L.check(lib.pda_conv3x3_tc(s0.data_ptr(), c0, ops._ptr(s1), c1, wp.data_ptr(), bias.data_ptr(),
                                   full.t.data_ptr(), pooled.t.data_ptr() if pool else 0, 0, B, H, W, cout, 1, 0, f16,
                                   ops.range_flag(dev).data_ptr() if f16 else 0, ops._stream()), "conv3x3")
        torch.cuda.synchronize()
    finally:
        lib.pda_set_conv_pair(prev)
    assert full.intact() and torch.equal(full.t, want_full)
    if pool:
        assert pooled.intact() and torch.equal(pooled.t, want_pool)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,h,w,c0,c1,cout", [(2, 17, 12, 128, 64, 64), (1, 9, 20, 256, 128, 128), (1, 8, 8, 512, 256, 256)])
def test_fused_upsample_conv_reads_and_writes_stay_inside(B, h, w, c0, c1, cout, dt):
    """The fused form reads the low-resolution tensor through TMA patches that overhang the image (zero fill) and writes
    the output by TMA: inputs and output all sit between canaries, the result equals upsample-then-conv bit for bit."""
    from probabilistic_domain_adaptation_b200 import ops
    lib, L = _lib()
    dev = _dev()
    g = torch.Generator().manual_seed(h * w + cout)
    low, bridge = Guarded((B, h, w, c0), dt, dev), Guarded((B, 2 * h, 2 * w, c1), dt, dev)
    low.t.copy_(torch.randn(B, h, w, c0, generator=g).to(dev).to(dt))
    bridge.t.copy_(torch.randn(B, 2 * h, 2 * w, c1, generator=g).to(dev).to(dt))
    wt = (torch.randn(cout, c0 + c1, 3, 3, generator=g) * 0.05).to(dev)
    bias = (torch.randn(cout, generator=g) * 0.1).to(dev)
    wp = ops.pack_conv3x3_weights(wt, dtype=dt)
    want, _ = ops.conv3x3(ops.upsample2x(low.t.contiguous()), bridge.t.contiguous(), wp, bias)
    out = Guarded((B, 2 * h, 2 * w, cout), dt, dev)
    f16 = int(dt == torch.float16)
    L.check(lib.pda_conv3x3_up_tc(low.t.data_ptr(), c0, bridge.t.data_ptr(), c1, wp.data_ptr(), bias.data_ptr(),
                                  out.t.data_ptr(), 0, B, 2 * h, 2 * w, cout, 1, f16,
                                  ops.range_flag(dev).data_ptr() if f16 else 0, ops._stream()), "conv3x3_up")
    torch.cuda.synchronize()
    assert out.intact() and low.intact() and bridge.intact()
    assert torch.equal(out.t, want)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_first_conv_upsample_pool_writes_stay_inside(dt):
    from probabilistic_domain_adaptation_b200 import ops
    lib, L = _lib()
    dev = _dev()
    g = torch.Generator().manual_seed(2)
    f16 = int(dt == torch.float16)
    B, H, W = 2, 10, 22
    x0 = torch.randn(B, 1, H, W, generator=g).to(dev)
    w = (torch.randn(64, 1, 3, 3, generator=g) * 0.3).to(dev)
    b = torch.zeros(64, device=dev)
    want = ops.conv3x3_first(x0, None, w, b, dtype=dt)
    out = Guarded((B, H, W, 64), dt, dev)
    L.check(lib.pda_conv3x3_first(x0.data_ptr(), 0, w.data_ptr(), b.data_ptr(), out.t.data_ptr(), B, H, W, 64, 1, f16,
                                  ops._stream()), "first")
    up_want = ops.upsample2x(want)
    up = Guarded((B, 2 * H, 2 * W, 64), dt, dev)
    L.check(lib.pda_upsample2x_bilinear(want.data_ptr(), up.t.data_ptr(), B, H, W, 64, f16, ops._stream()), "up")
    pl_want = ops.avgpool2(want)
    pl = Guarded((B, H // 2, W // 2, 64), dt, dev)
    L.check(lib.pda_avgpool2(want.data_ptr(), pl.t.data_ptr(), B, H, W, 64, f16, ops._stream()), "pool")
    torch.cuda.synchronize()
    assert out.intact() and torch.equal(out.t, want)
    assert up.intact() and torch.equal(up.t, up_want)
    assert pl.intact() and torch.equal(pl.t, pl_want)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,P_hw,S", [(2, (9, 13), 5), (1, (16, 24), 16), (3, (5, 5), 1)])
def test_fcomb_writes_stay_inside(B, P_hw, S, dt):
    from oracle import punet_oracle as po
    from probabilistic_domain_adaptation_b200 import ops
    lib, L = _lib()
    dev = _dev()
    h, w_ = P_hw
    sd = po.make_state_dict(0, last_layer_gain=8.0)
    k = ["fcomb.layers.0", "fcomb.layers.2", "fcomb.last_layer"]
    ws = [sd[f"{n}.{p}"].to(dev).contiguous() for n in k for p in ("weight", "bias")]
    g = torch.Generator().manual_seed(S)
    feat = torch.relu(torch.randn(B, h, w_, 64, generator=g)).to(dev).to(dt)
    z = torch.randn(S, B, 6, generator=g).to(dev)
    want = ops.fcomb_mc_consensus(feat, z, *ws, want_mean=True, want_weight=True, want_mask=True, want_logits=True,
                                  want_probs=True)
    P = h * w_
    mean, weight = Guarded((B, 1, h, w_), torch.float32, dev), Guarded((B, 1, h, w_), torch.float32, dev)
    mask = Guarded((B, 1, h, w_), torch.int64, dev)
    logits, probs = Guarded((S, B, 1, h, w_), torch.float32, dev), Guarded((S, B, 1, h, w_), torch.float32, dev)
    scratch = Guarded((int(lib.pda_fcomb_scratch_floats(S, B)),), torch.float32, dev)
    L.check(lib.pda_fcomb_mc_consensus(feat.data_ptr(), z.data_ptr(), *[t.data_ptr() for t in ws], B, P, S, 6, 0.9, 0.1,
                                       mean.t.data_ptr(), weight.t.data_ptr(), mask.t.data_ptr(), logits.t.data_ptr(),
                                       probs.t.data_ptr(), scratch.t.data_ptr(), int(dt == torch.float16),
                                       ops._stream()), "fcomb")
    torch.cuda.synchronize()
    for gd, key in ((mean, "mean"), (weight, "weight"), (mask, "mask"), (logits, "logits"), (probs, "probs")):
        assert gd.intact(), key
        assert torch.equal(gd.t, want[key]), key
    assert scratch.intact()


@pytest.mark.parametrize("B,H,W,c0,c1,cout", [(1, 18, 10, 64, 0, 64), (2, 34, 24, 128, 64, 128), (3, 16, 8, 64, 0, 256)])
def test_wgrad_writes_stay_inside(B, H, W, c0, c1, cout):
    from probabilistic_domain_adaptation_b200 import ops
    lib, L = _lib()
    dev = _dev()
    g = torch.Generator().manual_seed(cout + W)
    x = torch.randn(B, H, W, c0, generator=g).to(dev).to(torch.bfloat16)
    s1 = torch.randn(B, H, W, c1, generator=g).to(dev).to(torch.bfloat16) if c1 else None
    dz = torch.randn(B, H, W, cout, generator=g).to(dev).to(torch.bfloat16)
    want_dw, want_db = ops.conv3x3_wgrad(x, s1, dz, want_bias=True)
    ctot = c0 + c1
    scratch = Guarded((int(lib.pda_conv3x3_wgrad_scratch_floats(ctot, cout)),), torch.float32, dev)
    dw, db = Guarded((cout, ctot, 3, 3), torch.float32, dev), Guarded((cout,), torch.float32, dev)
    L.check(lib.pda_conv3x3_wgrad_bf16(x.data_ptr(), c0, ops._ptr(s1), c1, dz.data_ptr(), scratch.t.data_ptr(),
                                       dw.t.data_ptr(), db.t.data_ptr(), B, H, W, cout, 0, 0, ops._stream()), "wgrad")
    torch.cuda.synchronize()
    assert scratch.intact() and dw.intact() and db.intact()
    # scratch_is_zero = 1: the caller guarantees a zeroed scratch and the call leaves it zeroed again (twice in a row)
    scratch.t.zero_()
    for _ in range(2):
        dw2, db2 = torch.empty_like(want_dw), torch.empty_like(want_db)
        L.check(lib.pda_conv3x3_wgrad_bf16(x.data_ptr(), c0, ops._ptr(s1), c1, dz.data_ptr(), scratch.t.data_ptr(),
                                           dw2.data_ptr(), db2.data_ptr(), B, H, W, cout, 0, 1, ops._stream()), "wgrad")
        torch.cuda.synchronize()
        assert torch.allclose(dw2, want_dw, rtol=1e-4, atol=1e-4) and torch.allclose(db2, want_db, rtol=1e-4, atol=1e-3)
        assert not bool(scratch.t.any()), "the weight-gradient scratch must be all zero again after the launch"
    assert scratch.intact()
    # the flush uses fp32 atomics: equal up to summation order
    assert torch.allclose(dw.t, want_dw, rtol=1e-4, atol=1e-4) and torch.allclose(db.t, want_db, rtol=1e-4, atol=1e-3)


def test_repeated_launches_are_bit_identical():
    """Race check by repetition: 30 launches of the conv (pair and single), the fused Fcomb kernel and the tiled-forward
    path on the same inputs must give identical bits every time (none of them uses atomics)."""
    from oracle import punet_oracle as po
    from probabilistic_domain_adaptation_b200 import ProbabilisticUnet, consensus
    lib, _ = _lib()
    dev = _dev()
    m = ProbabilisticUnet(1, 1, [64, 128, 256, 512], 6, 3, 1.0).to(dev).eval()
    m.load_state_dict(po.make_state_dict(0, last_layer_gain=8.0))
    x, _, eps, _ = po.synthetic_inputs(2, 72, 104, s=8)
    x, eps = x.to(dev), eps.to(dev)
    ref = None
    for i in range(30):
        prev = lib.pda_set_conv_pair(i & 1)
        try:
            mean, mask = consensus.sample_from_teacher(m, x, 8, do_consensus_masking=True, eps=eps)
        finally:
            lib.pda_set_conv_pair(prev)
        if ref is None:
            ref = (mean.clone(), mask.clone())
        else:
            assert torch.equal(mean, ref[0]) and torch.equal(mask, ref[1]), i


def test_two_models_on_two_streams_run_the_whole_inference_concurrently():
    """Two models run their MC inference (all conv layers, the fused up-sampling convs, Fcomb) on two CUDA streams at the
    same time, at a size where every kernel fills the GPU, so that the CTAs of the two launches interleave: results must
    be bit-identical to the serial runs.  Regression test for the slab-ring protocol of the fused up-sampling conv: with a
    single "stage free" barrier shared by its two producers, a producer that reached a skipped use late waited for a
    completion only its own next chunk could cause -- dead-lock as soon as the kernel did not have the GPU to itself
    (bounded wait -> launch error after 2 s)."""
    from oracle import punet_oracle as po
    from probabilistic_domain_adaptation_b200 import ProbabilisticUnet, consensus
    dev = _dev()
    models = []
    for seed in (0, 1):
        m = ProbabilisticUnet(1, 1, [64, 128, 256, 512], 6, 3, 1.0).to(dev).eval()
        m.load_state_dict(po.make_state_dict(seed, last_layer_gain=8.0))
        models.append(m)
    x, _, eps, _ = po.synthetic_inputs(2, 512, 512, s=8)
    x, eps = x.to(dev), eps.to(dev)
    serial = [consensus.sample_from_teacher(m, x, 8, do_consensus_masking=True, eps=eps) for m in models]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for it in range(12):
        outs = []
        for m, st in zip(models, streams):
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                outs.append(consensus.sample_from_teacher(m, x, 8, do_consensus_masking=True, eps=eps))
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        for o, ref in zip(outs, serial):
            assert torch.equal(o[0], ref[0]) and torch.equal(o[1], ref[1]), it
