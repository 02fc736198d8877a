"""GPU parity of the conv / pool / upsample / head kernels (called through the C ABI) against a plain
PyTorch fp32 reference of the same op on the same bf16-rounded inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _ref_conv(srcs, w, bias, relu, pool):
    """srcs: list of NHWC bf16; w (cout, cin, 3, 3) fp32 (bf16-representable)."""
    x = torch.cat([s.float() for s in srcs], dim=3).permute(0, 3, 1, 2)
    y = F.conv2d(x.double(), w.double(), bias.double(), padding=1).float()
    if relu:
        y = F.relu(y)
    full = y.permute(0, 2, 3, 1).contiguous()
    pooled = F.avg_pool2d(y, 2).permute(0, 2, 3, 1).contiguous() if pool else None
    return full, pooled


CONV_CASES = [
    # B, H, W, c0, c1, cout, pool, bn
    (1, 16, 16, 64, 0, 64, False, 0),
    (2, 24, 40, 64, 0, 64, True, 0),
    (1, 8, 8, 128, 0, 128, True, 0),
    (1, 5, 9, 256, 0, 512, False, 0),
    (2, 16, 32, 128, 64, 64, False, 0),
    (1, 16, 16, 512, 256, 256, False, 256),
    (1, 32, 16, 256, 0, 256, True, 64),
    (1, 64, 64, 64, 0, 128, False, 128),
    # persistent kernel paths: two M-tiles per unit with ragged edges, concat with streamed weights, many units per CTA
    (1, 40, 24, 64, 64, 64, False, 0),
    (2, 34, 18, 128, 0, 128, True, 0),
    (1, 18, 10, 64, 0, 64, True, 0),
    (3, 48, 72, 64, 0, 64, False, 0),
    (1, 96, 160, 64, 0, 64, True, 0),
    (1, 17, 9, 128, 64, 256, False, 0),
]


# "tc": tensor-core kernel on bf16 activations (training format); "tc_f16": the same kernel on fp16 activations (the
# no-grad format, fp16 x fp16 MMAs); "simt": CUDA-core cross-check (bf16)
@pytest.mark.parametrize("B,H,W,c0,c1,cout,pool,bn", CONV_CASES)
@pytest.mark.parametrize("mode", ["tc", "tc_f16", "simt"])
def test_conv3x3_matches_fp32_reference(B, H, W, c0, c1, cout, pool, bn, mode):
    from probabilistic_domain_adaptation_b200 import ops
    dev = _dev()
    dt = torch.float16 if mode == "tc_f16" else torch.bfloat16
    ulp = 2 ** -11 if mode == "tc_f16" else 2 ** -8   # one ulp of the 16-bit output format (half an ulp + fp32 noise)
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + H * 10 + cout)
    s0 = torch.randn(B, H, W, c0, generator=g).to(dev).to(dt)
    s1 = torch.randn(B, H, W, c1, generator=g).to(dev).to(dt) if c1 else None
    ctot = c0 + c1
    w = (torch.randn(cout, ctot, 3, 3, generator=g) * (2.0 / (9 * ctot)) ** 0.5).to(dev)
    w = w.to(dt).float()  # representable in the operand format so that only accumulation order differs
    bias = (torch.randn(cout, generator=g) * 0.1).to(dev)
    wp = ops.pack_conv3x3_weights(w, dtype=dt)
    full, pooled = ops.conv3x3(s0, s1, wp, bias, relu=True, want_full=True, want_pool=pool, bn_tile=bn,
                               simt=(mode == "simt"))
    torch.cuda.synchronize()
    assert full.dtype == dt
    rfull, rpool = _ref_conv([s0] + ([s1] if c1 else []), w, bias, True, pool)
    err = (full.float() - rfull).abs()
    tol = ulp * rfull.abs() + 1e-3
    assert (err <= tol).all(), f"full: max err {err.max().item()} at {err.argmax().item()}"
    if pool:
        err = (pooled.float() - rpool).abs()
        tol = ulp * rpool.abs() + 1e-3
        assert (err <= tol).all(), f"pool: max err {err.max().item()}"


@pytest.mark.parametrize("B,H,W,c0,c1,cout,pool,bn", CONV_CASES + [(2, 64, 128, 64, 0, 64, True, 0),
                                                                      (1, 128, 136, 256, 128, 128, False, 0),
                                                                      (3, 24, 8, 64, 0, 256, False, 0)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_conv3x3_cta_pair_kernel_is_bit_identical(B, H, W, c0, c1, cout, pool, bn, dt):
    """csrc/conv3x3_tc2.cu (cta_group::2: one M = 256 MMA per CTA pair, weight tile split across the pair) against the
    single-CTA kernel: same operands, same fp32 accumulation per output row -> identical bits; odd tile counts exercise
    the ghost tile of the peer CTA."""
    from probabilistic_domain_adaptation_b200 import _lib, ops
    dev = _dev()
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(B * 77 + H + cout)
    s0 = torch.randn(B, H, W, c0, generator=g).to(dev).to(dt)
    s1 = torch.randn(B, H, W, c1, generator=g).to(dev).to(dt) if c1 else None
    ctot = c0 + c1
    w = (torch.randn(cout, ctot, 3, 3, generator=g) * (2.0 / (9 * ctot)) ** 0.5).to(dev)
    bias = (torch.randn(cout, generator=g) * 0.1).to(dev)
    wp = ops.pack_conv3x3_weights(w, dtype=dt)
    prev = lib.pda_set_conv_pair(0)
    try:
        f0, p0 = ops.conv3x3(s0, s1, wp, bias, relu=True, want_full=True, want_pool=pool, bn_tile=bn)
        lib.pda_set_conv_pair(1)
        f1, p1 = ops.conv3x3(s0, s1, wp, bias, relu=True, want_full=True, want_pool=pool, bn_tile=bn)
        # twice, so that a stale barrier phase / TMEM state of the first launch would show
        f2, p2 = ops.conv3x3(s0, s1, wp, bias, relu=True, want_full=True, want_pool=pool, bn_tile=bn)
        torch.cuda.synchronize()
    finally:
        lib.pda_set_conv_pair(prev)
    assert torch.equal(f0, f1) and torch.equal(f1, f2)
    if pool:
        assert torch.equal(p0, p1) and torch.equal(p1, p2)


def test_conv3x3_fp16_range_flag_and_saturation():
    """fp16 outputs saturate at +-65504 (never inf) and raise the sticky device flag; in-range launches leave it alone;
    the asynchronous poll switches the no-grad path to bf16."""
    import warnings
    from probabilistic_domain_adaptation_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(1)
    s0 = torch.rand(1, 16, 16, 64, generator=g).to(dev).to(torch.float16)
    w = (torch.rand(64, 64, 3, 3, generator=g) * 0.05).to(dev)
    bias = torch.zeros(64, device=dev)
    wp = ops.pack_conv3x3_weights(w, dtype=torch.float16)
    flag = ops.range_flag(dev)
    flag.zero_()
    full, _ = ops.conv3x3(s0, None, wp, bias)
    assert int(flag.item()) == 0 and torch.isfinite(full).all()
    big, _ = ops.conv3x3((s0.float() * 6e4).clamp(max=6e4).to(torch.float16), None, wp, bias)
    assert int(flag.item()) == 1
    assert torch.isfinite(big).all() and float(big.float().max()) == 65504.0
    saved = ops.INFER_DTYPE
    try:
        ops.INFER_DTYPE = torch.float16
        with warnings.catch_warnings(record=True) as rec:
            warnings.simplefilter("always")
            assert ops.check_fp16_range(dev) is False
        assert ops.INFER_DTYPE == torch.bfloat16 and int(flag.item()) == 0
        assert any("fp16 range" in str(r.message) for r in rec)
    finally:
        ops.INFER_DTYPE = saved


def test_conv3x3_pool_only_output():
    from probabilistic_domain_adaptation_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(7)
    s0 = torch.randn(1, 16, 16, 64, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(64, 64, 3, 3, generator=g) * 0.06).to(dev).to(torch.bfloat16).float()
    bias = torch.zeros(64, device=dev)
    wp = ops.pack_conv3x3_weights(w)
    f1, p1 = ops.conv3x3(s0, None, wp, bias, want_full=True, want_pool=True)
    f2, p2 = ops.conv3x3(s0, None, wp, bias, want_full=False, want_pool=True)
    assert f2 is None and torch.equal(p1, p2)


DTYPES = [torch.bfloat16, torch.float16]
ULP = {torch.bfloat16: 2 ** -8, torch.float16: 2 ** -11}


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("cin", [1, 2])
def test_first_conv(cin, dt):
    from probabilistic_domain_adaptation_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, cin, 24, 40, generator=g).to(dev)
    w = (torch.randn(64, cin, 3, 3, generator=g) * 0.4).to(dev)
    b = (torch.randn(64, generator=g) * 0.1).to(dev)
    out = ops.conv3x3_first(x[:, 0:1].contiguous(), x[:, 1:2].contiguous() if cin == 2 else None, w, b, dtype=dt)
    assert out.dtype == dt
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1)).float().permute(0, 2, 3, 1)
    err = (out.float() - ref).abs()
    assert (err <= ULP[dt] * ref.abs() + 1e-5).all(), err.max().item()


@pytest.mark.parametrize("dt", DTYPES)
def test_avgpool_and_upsample(dt):
    from probabilistic_domain_adaptation_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 10, 18, 128, generator=g).to(dev).to(dt)
    xn = x.float().permute(0, 3, 1, 2)
    p = ops.avgpool2(x)
    rp = F.avg_pool2d(xn, 2, 2, 0, ceil_mode=True).permute(0, 2, 3, 1)
    assert p.dtype == dt and ((p.float() - rp).abs() <= ULP[dt] * rp.abs() + 1e-6).all()
    u = ops.upsample2x(x)
    ru = F.interpolate(xn, mode="bilinear", scale_factor=2, align_corners=True).permute(0, 2, 3, 1)
    err = (u.float() - ru).abs()
    assert u.dtype == dt and (err <= ULP[dt] * ru.abs() + 1e-5).all(), err.max().item()


@pytest.mark.parametrize("dt", DTYPES)
def test_gauss_head_and_latents(dt):
    from probabilistic_domain_adaptation_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(9)
    enc = torch.randn(3, 5, 9, 512, generator=g).to(dev).to(dt)
    w = (torch.randn(12, 512, 1, 1, generator=g) * 0.05).to(dev)
    b = (torch.randn(12, generator=g) * 0.01).to(dev)
    out = ops.gauss_head(enc, w, b, 6)
    e = enc.double().permute(0, 3, 1, 2)  # double: cuDNN fp32 conv may run in TF32
    e = torch.mean(torch.mean(e, dim=2, keepdim=True), dim=3, keepdim=True)
    ref = F.conv2d(e, w.double(), b.double())[:, :, 0, 0].float()
    assert torch.allclose(out, ref, atol=1e-5), (out - ref).abs().max().item()
    eps = torch.randn(4, 3, 6, generator=g).to(dev)
    z = ops.latent_samples(out, eps)
    assert torch.allclose(z, out[None, :, :6] + torch.exp(out[None, :, 6:]) * eps, atol=1e-6)
    q = out.clone()
    q[:, :6] += 0.3
    kl = ops.kl_diag_gauss(q, out)
    d = torch.distributions
    ref_kl = d.kl.kl_divergence(d.Independent(d.Normal(q[:, :6], torch.exp(q[:, 6:])), 1),
                                d.Independent(d.Normal(out[:, :6], torch.exp(out[:, 6:])), 1))
    assert torch.allclose(kl, ref_kl, rtol=1e-5, atol=1e-6)


def test_multi_tensor_ema_bit_exact():
    from probabilistic_domain_adaptation_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(11)
    shapes = [(64, 1, 3, 3), (64,), (512, 256, 3, 3), (12, 512, 1, 1), (7,), (70001,)]
    teacher = [torch.randn(s, generator=g).to(dev) for s in shapes]
    student = [torch.randn(s, generator=g).to(dev) for s in shapes]
    m = 0.999
    ref = [t * m + p * (1. - m) for t, p in zip(teacher, student)]  # mean_teacher_trainer.py:55
    table = ops.build_ema_table(teacher, student)
    ops.multi_tensor_ema(table, m)
    for t, r in zip(teacher, ref):
        assert torch.equal(t, r)


def test_refresh_packed_matches_single_tensor_pack():
    """autograd_ops.refresh_packed (ONE launch for all convs of a module: forward bf16, rot180 bf16, forward fp16) equals
    the per-conv pack kernel bit for bit."""
    import torch.nn as nn
    from probabilistic_domain_adaptation_b200 import autograd_ops, ops
    dev = _dev()
    torch.manual_seed(3)
    mod = nn.Sequential(nn.Conv2d(64, 128, 3, padding=1), nn.Conv2d(128, 64, 3, padding=1),
                        nn.Conv2d(192, 256, 3, padding=1), nn.Conv2d(1, 64, 3, padding=1)).to(dev)
    autograd_ops.refresh_packed(mod, rot180=True, bf16=True, f16=True)
    for conv in list(mod)[:3]:
        w = conv.weight.detach()
        assert torch.equal(conv.__dict__["_pda_packed"][1], ops.pack_conv3x3_weights(w))
        assert torch.equal(conv.__dict__["_pda_packed_rot"][1], ops.pack_conv3x3_weights(w, rot180=True))
        if ops.INFER_DTYPE == torch.float16:
            assert torch.equal(conv.__dict__["_pda_packed_f16"][1], ops.pack_conv3x3_weights(w, dtype=torch.float16))
    assert "_pda_packed" not in mod[3].__dict__          # the cin = 1 first layer is not a tensor-core conv


@pytest.mark.parametrize("budget", [148, 4])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,h,w,c0,c1,cout", [(1, 16, 16, 128, 64, 64), (2, 20, 12, 256, 128, 128), (1, 8, 24, 512, 256, 256),
                                             (3, 9, 4, 128, 64, 64), (1, 64, 64, 128, 64, 64), (2, 33, 17, 256, 128, 128),
                                             (1, 16, 16, 512, 256, 256), (3, 5, 9, 64, 64, 128), (1, 100, 36, 128, 64, 128)])
def test_fused_upsample_conv_is_bit_identical_to_upsample_then_conv(B, h, w, c0, c1, cout, dt, budget):
    """pda_conv3x3_up_tc (the epilogue warps interpolate the first K segment into the operand slabs) against
    pda_upsample2x_bilinear + pda_conv3x3_tc: the same bits, on ragged tiles, odd tile counts (ghost tile of the peer CTA),
    repeated launches, and with the grid cut to two CTA pairs (budget 4) so that every pair walks MANY units -- the ring
    phases, the patch prefetch across unit boundaries and the produce / drain interleave only show up there (a first
    version passed every single-unit shape and dead-locked at 1024^2)."""
    from probabilistic_domain_adaptation_b200 import _lib, ops
    lib = _lib.load()
    dev = _dev()
    g = torch.Generator().manual_seed(B * 31 + h + c0)
    x_low = torch.randn(B, h, w, c0, generator=g).to(dev).to(dt)
    bridge = torch.randn(B, 2 * h, 2 * w, c1, generator=g).to(dev).to(dt)
    wt = (torch.randn(cout, c0 + c1, 3, 3, generator=g) * (2.0 / (9 * (c0 + c1))) ** 0.5).to(dev)
    bias = (torch.randn(cout, generator=g) * 0.1).to(dev)
    wp = ops.pack_conv3x3_weights(wt, dtype=dt)
    if not ops.can_fuse_upsample(x_low, bridge):
        pytest.skip("single-tile shape: the fused form needs the CTA-pair kernel")
    want, _ = ops.conv3x3(ops.upsample2x(x_low), bridge, wp, bias)
    prev = lib.pda_set_sm_budget(budget)
    try:
        for _ in range(3):
            got = ops.conv3x3_up(x_low, bridge, wp, bias)
            torch.cuda.synchronize()
            assert torch.equal(got, want)
    finally:
        lib.pda_set_sm_budget(prev if prev > 0 else 148)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,H,W,cin", [(1, 5, 7, 1), (2, 37, 129, 2), (3, 16, 8, 1), (1, 130, 131, 2), (4, 64, 64, 1)])
def test_first_conv_tensor_core_kernel_matches_cuda_core_kernel(B, H, W, cin, dt):
    """The first layer on the tensor core (kind::tf32 MMA with hi / lo split operands, one pixel per TMEM lane) against
    the CUDA-core kernel that does the same conv with fp32 FMAs: equal up to the summation order in fp32 and the final
    16-bit rounding -- on images smaller than one 128-pixel tile, widths that are not a multiple of anything, several
    images per tile, and twice in a row (barrier phases / TMEM state of a previous launch)."""
    from probabilistic_domain_adaptation_b200 import _lib, ops
    lib = _lib.load()
    dev = _dev()
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + cin)
    x0 = (torch.randn(B, 1, H, W, generator=g) * 3.0).to(dev)
    x1 = (torch.randn(B, 1, H, W, generator=g) * 3.0).to(dev) if cin == 2 else None
    w = (torch.randn(64, cin, 3, 3, generator=g) * 0.4).to(dev)
    b = torch.randn(64, generator=g).to(dev)
    prev = lib.pda_set_first_conv_tc(0)
    try:
        want = ops.conv3x3_first(x0, x1, w, b, dtype=dt)
        lib.pda_set_first_conv_tc(1)
        for _ in range(2):
            got = ops.conv3x3_first(x0, x1, w, b, dtype=dt)
            torch.cuda.synchronize()
            ulp = 2.0 ** -7 if dt == torch.bfloat16 else 2.0 ** -10   # spacing of the 16-bit format relative to the value
            err = (got.float() - want.float()).abs()
            # at most one unit in the last place of the 16-bit result; near the ReLU threshold the split-operand sum may
            # differ by ~1e-5 absolute (2^-21 of the products' magnitude)
            assert bool((err <= ulp * want.float().abs() + 1e-4).all()), float(err.max())
            assert float((got != want).float().mean()) < 0.02     # and that only where a sum sits on a rounding boundary
    finally:
        lib.pda_set_first_conv_tc(prev)
