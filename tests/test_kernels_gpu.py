"""Per-kernel numerics: every C-ABI entry point against a plain torch fp32 reference of the same op
on the same seeded inputs (the 16-bit-rounded inputs are upcast, so only accumulation order and the
16-bit output rounding differ).  Tolerances are stated per test for fp16 storage; every test runs a
second time against the bf16-storage build of the library (libidiff_b200_bf16.so) with the tolerances
scaled by 8 = 2^(11-8), the ratio of the two formats' rounding steps.
"""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from instancediffusion_b200 import ops
    return ops


TOL_SCALE = 1.0


@pytest.fixture(autouse=True, params=["fp16", "bf16"])
def storage(request):
    """Run every test once per storage type of the library."""
    global TOL_SCALE
    ops = _ops()
    bf = request.param == "bf16"
    ops.set_storage_dtype(torch.bfloat16 if bf else torch.float16)
    TOL_SCALE = 8.0 if bf else 1.0
    yield request.param
    ops.set_storage_dtype(torch.float16)
    TOL_SCALE = 1.0


def _h(t):
    """To the current 16-bit storage type."""
    return t.to(_ops().HALF)


def _randn(shape, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev)


def _check(got, ref, rtol, atol, what):
    got = got.float()
    ref = ref.float()
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    err = (got - ref).abs()
    tol = TOL_SCALE * (atol + rtol * ref.abs())
    bad = (err > tol)
    max_err = err.max().item()
    ref_max = ref.abs().max().item()
    print(f"[{what}] max_abs_err={max_err:.3e} ref_max={ref_max:.3e} bad={int(bad.sum())}/{bad.numel()}")
    assert not bad.any(), f"{what}: max err {max_err:.3e} (ref max {ref_max:.3e}), {int(bad.sum())} bad"


# --------------------------------------------------------------------------------------------
# GEMM (linear)
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 320, 320), (4096, 640, 320), (100, 320, 1280),
                                   (8, 1280, 320), (1024, 1280, 5120), (77, 640, 768)])
def test_gemm_linear(cuda_device, M, N, K):
    ops = _ops()
    a = _h(_randn((M, K), cuda_device, 1.0, 1))
    w = _h(_randn((N, K), cuda_device, 1.0 / math.sqrt(K), 2))
    bias = _randn((N,), cuda_device, 0.5, 3)
    out = ops.gemm(a, w, bias)
    ref = a.float() @ w.float().t() + bias
    # fp32 accumulate, fp16 output rounding: 1e-3 relative + small absolute
    _check(out, ref, 2e-3, 2e-3, f"gemm {M}x{N}x{K}")


@pytest.mark.parametrize("M,N,K", [(512, 1280, 11520), (2048, 1280, 5120), (8192, 640, 2560), (2048, 3840, 1280),
                                   (512, 320, 23040), (19000, 960, 640)])
def test_gemm_stream_k_shapes(cuda_device, M, N, K):
    """Tile counts that are not multiples of the SM count: the ragged waves are split over all SMs in
    k-block units and fixed up through the fp32 workspace; the result must be bit-reproducible."""
    ops = _ops()
    a = _h(_randn((M, K), cuda_device, 1.0, 1))
    w = _h(_randn((N, K), cuda_device, 1.0 / math.sqrt(K), 2))
    bias = _randn((N,), cuda_device, 0.5, 3)
    res = _h(_randn((M, N), cuda_device, 1.0, 4))
    ref = res.float() + 0.5 * (a.float() @ w.float().t() + bias)
    # the planner may prefer plain rounds for a shape; IDIFF_GEMM_PLAN="0,1" keeps its tile width and
    # forces the stream-K schedule, so both paths are covered
    for plan in (None, "0,1"):
        if plan is None:
            os.environ.pop("IDIFF_GEMM_PLAN", None)
        else:
            os.environ["IDIFF_GEMM_PLAN"] = plan
        try:
            out = ops.gemm(a, w, bias, residual=res, gate=0.5)
            _check(out, ref, 2e-3, 2e-3, f"gemm stream-K {M}x{N}x{K} plan={plan}")
            for _ in range(3):
                again = ops.gemm(a, w, bias, residual=res, gate=0.5)
                assert torch.equal(out, again), "stream-K result is not bit-reproducible"
        finally:
            os.environ.pop("IDIFF_GEMM_PLAN", None)


def test_gemm_geglu_stream_k(cuda_device):
    from instancediffusion_b200.packing import pack_geglu
    ops = _ops()
    M, C = 2048, 1280
    a = _h(_randn((M, C), cuda_device, 1.0, 1))
    w = _h(_randn((8 * C, C), cuda_device, 1.0 / math.sqrt(C), 2))
    bias = _randn((8 * C,), cuda_device, 0.5, 3)
    wp, bp = pack_geglu(w, bias)
    h = a.float() @ w.float().t() + bias
    x, gate = h.chunk(2, dim=-1)
    os.environ["IDIFF_GEMM_PLAN"] = "0,1"  # force the stream-K schedule (see above)
    try:
        out = ops.gemm(a, wp, bp, geglu=True)
        _check(out, x * F.gelu(gate), 3e-3, 3e-3, "geglu stream-K 2048x1280")
        assert torch.equal(out, ops.gemm(a, wp, bp, geglu=True))
    finally:
        os.environ.pop("IDIFF_GEMM_PLAN", None)
    _check(ops.gemm(a, wp, bp, geglu=True), x * F.gelu(gate), 3e-3, 3e-3, "geglu planned 2048x1280")


def test_gemm_residual_gate_silu(cuda_device):
    ops = _ops()
    M, N, K = 512, 320, 640
    a = _h(_randn((M, K), cuda_device, 1.0, 1))
    w = _h(_randn((N, K), cuda_device, 1.0 / math.sqrt(K), 2))
    bias = _randn((N,), cuda_device, 0.5, 3)
    res = _h(_randn((M, N), cuda_device, 1.0, 4))
    out = ops.gemm(a, w, bias, residual=res, gate=0.37)
    ref = res.float() + 0.37 * (a.float() @ w.float().t() + bias)
    _check(out, ref, 2e-3, 2e-3, "gemm residual+gate")
    out = ops.gemm(a, w, bias, silu=True)
    ref = F.silu(a.float() @ w.float().t() + bias)
    _check(out, ref, 2e-3, 2e-3, "gemm silu")
    # per-batch row vector (ResBlock emb add), 4 batches of 128 rows
    radd = _h(_randn((4, N), cuda_device, 1.0, 5))
    out = ops.gemm(a, w, bias, rowadd=radd, rows_per_batch=128)
    ref = a.float() @ w.float().t() + bias + radd.float().repeat_interleave(128, dim=0)
    _check(out, ref, 2e-3, 2e-3, "gemm rowadd")


@pytest.mark.parametrize("M,C", [(256, 320), (1000, 640)])
def test_gemm_geglu(cuda_device, M, C):
    """GEGLU (attention.py:36-43): proj -> chunk(2) -> x * gelu(gate); weight rows are packed per 64."""
    from instancediffusion_b200.packing import pack_geglu
    ops = _ops()
    inner = 4 * C
    a = _h(_randn((M, C), cuda_device, 1.0, 1))
    w = _h(_randn((2 * inner, C), cuda_device, 1.0 / math.sqrt(C), 2))
    bias = _randn((2 * inner,), cuda_device, 0.5, 3)
    wp, bp = pack_geglu(w, bias)
    out = ops.gemm(a, wp, bp, geglu=True)
    h = a.float() @ w.float().t() + bias
    x, gate = h.chunk(2, dim=-1)
    ref = x * F.gelu(gate)
    _check(out, ref, 3e-3, 3e-3, f"geglu {M}x{C}")


# --------------------------------------------------------------------------------------------
# conv3x3 (implicit GEMM through 4-D TMA)
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 64, 64, 64, 128), (2, 32, 32, 320, 320), (2, 16, 16, 640, 1280),
                                            (4, 8, 8, 1280, 1280), (3, 8, 8, 128, 64), (1, 24, 24, 64, 64),
                                            (2, 12, 12, 64, 64), (1, 96, 96, 64, 64)])
def test_conv3x3(cuda_device, B, H, W, Cin, Cout):
    from instancediffusion_b200.packing import pack_conv3x3
    ops = _ops()
    x = _h(_randn((B, Cin, H, W), cuda_device, 1.0, 1))
    w = _h(_randn((Cout, Cin, 3, 3), cuda_device, 1.0 / math.sqrt(9 * Cin), 2))
    bias = _randn((Cout,), cuda_device, 0.5, 3)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().reshape(B * H * W, Cin)
    out = ops.gemm(x_nhwc, pack_conv3x3(w), bias, conv=(B, H, W, Cin))
    ref = F.conv2d(x.float(), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(B * H * W, Cout)
    _check(out, ref, 2e-3, 2e-3, f"conv3x3 {B}x{H}x{W} {Cin}->{Cout}")


def test_conv3x3_rowadd_and_nchw_out(cuda_device):
    from instancediffusion_b200.packing import pack_conv3x3
    ops = _ops()
    B, H, W, Cin, Cout = 2, 64, 64, 320, 320
    x = _h(_randn((B, Cin, H, W), cuda_device, 1.0, 1))
    w = _h(_randn((Cout, Cin, 3, 3), cuda_device, 1.0 / math.sqrt(9 * Cin), 2))
    bias = _randn((Cout,), cuda_device, 0.5, 3)
    emb = _h(_randn((B, Cout), cuda_device, 1.0, 4))
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().reshape(B * H * W, Cin)
    out = ops.gemm(x_nhwc, pack_conv3x3(w), bias, conv=(B, H, W, Cin), rowadd=emb)
    ref = F.conv2d(x.float(), w.float(), bias, padding=1) + emb.float()[:, :, None, None]
    _check(out, ref.permute(0, 2, 3, 1).reshape(B * H * W, Cout), 2e-3, 2e-3, "conv3x3 + emb")
    # final conv 320 -> 4 with fp32 NCHW output (openaimodel.py:461-465)
    w4 = _h(_randn((4, Cin, 3, 3), cuda_device, 1.0 / math.sqrt(9 * Cin), 5))
    b4 = _randn((4,), cuda_device, 0.5, 6)
    w4p = torch.zeros((8, 9 * Cin), dtype=_ops().HALF, device=cuda_device)
    w4p[:4] = pack_conv3x3(w4)
    o = torch.empty((B, 4, H, W), dtype=torch.float32, device=cuda_device)
    ops.gemm(x_nhwc, w4p[:4], b4, conv=(B, H, W, Cin), out_nchw=o)
    ref = F.conv2d(x.float(), w4.float(), b4, padding=1)
    _check(o, ref, 1e-3, 1e-3, "conv3x3 -> fp32 NCHW")


def test_downsample_and_upsample_helpers(cuda_device):
    from instancediffusion_b200.packing import pack_conv3x3
    ops = _ops()
    B, H, W, C = 2, 32, 32, 320
    x = _h(_randn((B, C, H, W), cuda_device, 1.0, 1))
    w = _h(_randn((C, C, 3, 3), cuda_device, 1.0 / math.sqrt(9 * C), 2))
    bias = _randn((C,), cuda_device, 0.5, 3)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().reshape(B * H * W, C)
    cols = ops.im2col_s2(x_nhwc, B, H, W)
    out = ops.gemm(cols, pack_conv3x3(w), bias)
    ref = F.conv2d(x.float(), w.float(), bias, stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, C)
    _check(out, ref, 2e-3, 2e-3, "downsample conv (im2col s2)")
    up = ops.upsample_nearest2x(x_nhwc, B, H, W)
    ref = F.interpolate(x.float(), scale_factor=2, mode="nearest").permute(0, 2, 3, 1).reshape(-1, C)
    _check(up, ref, 0, 0, "upsample nearest 2x")


# --------------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------------
def _ref_attn(q, k, v, heads, scale):
    B, Nq, C = q.shape
    d = C // heads
    qh = q.float().view(B, Nq, heads, d).transpose(1, 2)
    kh = k.float().view(B, -1, heads, d).transpose(1, 2)
    vh = v.float().view(B, -1, heads, d).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * scale
    o = s.softmax(-1) @ vh
    return o.transpose(1, 2).reshape(B * Nq, C)


@pytest.mark.parametrize("B,N,d", [(2, 256, 40), (1, 4096, 40), (2, 1024, 80), (3, 256, 160), (4, 64, 160), (1, 200, 80)])
def test_self_attention_fused_qkv(cuda_device, B, N, d):
    ops = _ops()
    heads = 8
    C = heads * d
    qkv = _h(_randn((B * N, 3 * C), cuda_device, 1.0, 1))
    out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], batch=B, heads=heads, head_dim=d,
                        nq=N, n0=N, scale=d ** -0.5)
    q, k, v = (qkv[:, i * C:(i + 1) * C].reshape(B, N, C) for i in range(3))
    ref = _ref_attn(q, k, v, heads, d ** -0.5)
    # P is rounded to fp16 before the PV product (as in flash attention); output fp16
    _check(out, ref, 3e-3, 3e-3, f"self-attn B{B} N{N} d{d}")


@pytest.mark.parametrize("B,N,d,shared", [(2, 256, 40, False), (2, 1024, 80, True), (2, 64, 160, False), (1, 4096, 40, True)])
def test_gated_attention_two_segments(cuda_device, B, N, d, shared):
    """Keys = N visual tokens + 184 object tokens, queries = visual rows only (attention.py:304-308)."""
    ops = _ops()
    heads = 8
    C = heads * d
    qkv = _h(_randn((B * N, 3 * C), cuda_device, 1.0, 1))
    B1 = 1 if shared else B
    okv = _h(_randn((B1 * 184, 2 * C), cuda_device, 1.0, 2))
    out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], batch=B, heads=heads, head_dim=d,
                        nq=N, n0=N, scale=d ** -0.5, k1=okv[:, :C], v1=okv[:, C:], n1=184, kv1_batch=B1)
    q, k, v = (qkv[:, i * C:(i + 1) * C].reshape(B, N, C) for i in range(3))
    ok = okv[:, :C].reshape(B1, 184, C).expand(B, 184, C)
    ov = okv[:, C:].reshape(B1, 184, C).expand(B, 184, C)
    ref = _ref_attn(q, torch.cat([k, ok], 1), torch.cat([v, ov], 1), heads, d ** -0.5)
    _check(out, ref, 3e-3, 3e-3, f"gated-attn B{B} N{N} d{d} shared={shared}")


@pytest.mark.parametrize("B,N,d", [(2, 1024, 80), (2, 4096, 40), (2, 64, 160)])
def test_cross_attention_77_keys(cuda_device, B, N, d):
    ops = _ops()
    heads = 8
    C = heads * d
    q = _h(_randn((B * N, C), cuda_device, 1.0, 1))
    kv = _h(_randn((B * 77, 2 * C), cuda_device, 1.0, 2))
    out = ops.attention(q, kv[:, :C], kv[:, C:], batch=B, heads=heads, head_dim=d, nq=N, n0=77, scale=d ** -0.5)
    ref = _ref_attn(q.reshape(B, N, C), kv[:, :C].reshape(B, 77, C), kv[:, C:].reshape(B, 77, C), heads, d ** -0.5)
    _check(out, ref, 3e-3, 3e-3, f"cross-attn B{B} N{N} d{d}")


# --------------------------------------------------------------------------------------------
# normalisation
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,HW,C,silu,eps", [(2, 4096, 320, True, 1e-5), (2, 1024, 960, True, 1e-5), (3, 64, 2560, True, 1e-5),
                                             (2, 256, 1280, False, 1e-6), (1, 1024, 1920, True, 1e-5)])
@pytest.mark.parametrize("two_kernel", [False, True])
def test_groupnorm(cuda_device, B, HW, C, silu, eps, two_kernel):
    """Default dispatch (single-pass cluster kernel where the launch fits one wave) and the forced two-kernel path."""
    ops = _ops()
    x = _h((_randn((B * HW, C), cuda_device, 1.5, 1) + 0.3))
    gamma = 1.0 + _randn((C,), cuda_device, 0.2, 2)
    beta = _randn((C,), cuda_device, 0.2, 3)
    if two_kernel:
        os.environ["IDIFF_GN_FUSED"] = "0"
    try:
        out = ops.groupnorm(x, gamma, beta, batch=B, hw=HW, groups=32, eps=eps, silu=silu)
    finally:
        os.environ.pop("IDIFF_GN_FUSED", None)
    xr = x.float().view(B, HW, C).permute(0, 2, 1)
    ref = F.group_norm(xr, 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1).reshape(B * HW, C)
    _check(out, ref, 2e-3, 2e-3, f"groupnorm B{B} HW{HW} C{C}")


@pytest.mark.parametrize("B,HW,C", [(8, 4096, 320), (8, 4096, 640), (8, 4096, 960), (8, 1024, 640), (8, 1024, 1280),
                                    (8, 1024, 1920), (8, 1024, 960), (8, 256, 1280), (8, 256, 2560), (8, 256, 1920),
                                    (8, 64, 1280), (8, 64, 2560), (2, 4096, 320), (1, 256, 1280)])
def test_groupnorm_single_pass(cuda_device, B, HW, C):
    """The single-pass cluster kernel (one pass, statistics exchanged through distributed shared memory; the
    default for shapes that fit one wave, IDIFF_GN_FUSED=0 forces the two-kernel path) at every GroupNorm
    shape of the UNet; it must agree with torch and, to fp32 summation-order noise, with the two-kernel path."""
    ops = _ops()
    x = _h((_randn((B * HW, C), cuda_device, 1.5, 1) + 0.3))
    gamma = 1.0 + _randn((C,), cuda_device, 0.2, 2)
    beta = _randn((C,), cuda_device, 0.2, 3)
    os.environ["IDIFF_GN_FUSED"] = "0"
    try:
        base = ops.groupnorm(x, gamma, beta, batch=B, hw=HW, groups=32, eps=1e-5, silu=True)
    finally:
        os.environ.pop("IDIFF_GN_FUSED", None)
    out = ops.groupnorm(x, gamma, beta, batch=B, hw=HW, groups=32, eps=1e-5, silu=True)
    again = ops.groupnorm(x, gamma, beta, batch=B, hw=HW, groups=32, eps=1e-5, silu=True)
    ref = F.silu(F.group_norm(x.float().view(B, HW, C).permute(0, 2, 1), 32, gamma, beta, 1e-5))
    ref = ref.permute(0, 2, 1).reshape(B * HW, C)
    _check(out, ref, 2e-3, 2e-3, f"single-pass groupnorm B{B} HW{HW} C{C}")
    _check(out, base.float(), 2e-3, 2e-3, f"single-pass vs two-kernel groupnorm B{B} HW{HW} C{C}")
    assert torch.equal(out, again), "single-pass groupnorm is not deterministic"


@pytest.mark.parametrize("rows,C", [(4096, 320), (1000, 640), (77, 1280)])
def test_layernorm(cuda_device, rows, C):
    ops = _ops()
    x = _h((_randn((rows, C), cuda_device, 1.5, 1) + 0.3))
    gamma = 1.0 + _randn((C,), cuda_device, 0.2, 2)
    beta = _randn((C,), cuda_device, 0.2, 3)
    out = ops.layernorm(x, gamma, beta, 1e-5)
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    _check(out, ref, 2e-3, 2e-3, f"layernorm {rows}x{C}")


# --------------------------------------------------------------------------------------------
# ScaleU / Fourier filter
# --------------------------------------------------------------------------------------------
def _fourier_filter_ref(x, threshold, scale):
    """Restatement of openaimodel.py:25-48 in fp32/complex64."""
    import torch.fft as fft
    B, C, H, W = x.shape
    xf = fft.fftshift(fft.fftn(x.float(), dim=(-2, -1)), dim=(-2, -1))
    mask = torch.ones((B, C, H, W), device=x.device)
    crow, ccol = H // 2, W // 2
    mask[..., crow - threshold:crow + threshold, ccol - threshold:ccol + threshold] = scale
    xf = fft.ifftshift(xf * mask, dim=(-2, -1))
    return fft.ifftn(xf, dim=(-2, -1)).real


@pytest.mark.parametrize("B,H,W,C1,C2", [(2, 8, 8, 1280, 1280), (2, 16, 16, 1280, 640), (1, 64, 64, 320, 320), (1, 24, 24, 64, 64)])
def test_scaleu_concat(cuda_device, B, H, W, C1, C2):
    ops = _ops()
    h = _h(_randn((B, C1, H, W), cuda_device, 1.0, 1))
    skip = _h((_randn((B, C2, H, W), cuda_device, 1.0, 2) + 0.5))
    b_param = _randn((C1,), cuda_device, 0.5, 3)
    s_param = 0.4
    b1 = torch.tanh(b_param) + 1
    s = math.tanh(s_param) + 1
    hn = h.permute(0, 2, 3, 1).reshape(B * H * W, C1).contiguous()
    sn = skip.permute(0, 2, 3, 1).reshape(B * H * W, C2).contiguous()
    out = ops.scaleu_concat(hn, sn, b1, s, batch=B, height=H, width=W)
    ref_h = h.float() * b1[None, :, None, None]
    ref_s = _fourier_filter_ref(skip, 1, s)
    ref = torch.cat([ref_h, ref_s], 1).permute(0, 2, 3, 1).reshape(B * H * W, C1 + C2)
    _check(out, ref, 2e-3, 2e-3, f"scaleu {B}x{H}x{W} {C1}+{C2}")


# --------------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------------
def test_layout_roundtrip(cuda_device):
    ops = _ops()
    x = _randn((3, 4, 64, 64), cuda_device, 1.0, 1)
    y = ops.nchw_f32_to_nhwc_f16(x, 64)
    ref = torch.zeros((3 * 4096, 64), device=cuda_device)
    ref[:, :4] = x.permute(0, 2, 3, 1).reshape(-1, 4)
    _check(y, _h(ref), 0, 0, "nchw->nhwc pad")
    z = _h(_randn((2 * 256, 320), cuda_device, 1.0, 2))
    back = ops.nhwc_f16_to_nchw_f32(z, 2, 16, 16)
    _check(back, z.float().view(2, 256, 320).permute(0, 2, 1).reshape(2, 320, 16, 16), 0, 0, "nhwc->nchw")


def test_timestep_embedding(cuda_device):
    ops = _ops()
    t = torch.tensor([981.0, 1.0, 501.0, 21.0], device=cuda_device)
    out = ops.timestep_embedding(t, 320)
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=cuda_device) / half)
    args = t[:, None] * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1)
    _check(out, ref, 0, 1.5e-3, "timestep_embedding")


@pytest.mark.parametrize("D,mode,dropped", [(4, 0, False), (2, 0, True), (40, 1, False), (512, 1, False)])
def test_fourier_embed(cuda_device, D, mode, dropped):
    ops = _ops()
    rows = 60
    coords = torch.rand((rows, D), generator=torch.Generator().manual_seed(1)).to(cuda_device)
    masks = (torch.arange(rows) % 3 != 0).float().to(cuda_device)
    if mode == 1:
        coords[masks == 0] = 0
        coords[5] = 0.1  # padded slot with non-zero coords still counts as present (:267)
    text = _randn((rows, 768), cuda_device, 1.0, 2)
    null_text = _randn((768,), cuda_device, 1.0, 3)
    null_pos = _randn((32 * D,), cuda_device, 1.0, 4)
    out = torch.empty((rows, 768 + 32 * D), dtype=_ops().HALF, device=cuda_device)
    ops.fourier_embed(coords, masks, null_pos, out, text=text, null_text=null_text, mask_mode=mode, dropped=dropped)
    freqs = 100 ** (torch.arange(16, device=cuda_device) / 16)
    emb = torch.cat([f(fr * coords) for fr in freqs for f in (torch.sin, torch.cos)], -1)
    m = masks[:, None]
    if dropped:
        mp = torch.zeros_like(m)
    elif mode == 0:
        mp = m
    else:
        mp = ((coords.sum(-1, keepdim=True) + m) > 0).float()
    ref = torch.cat([text * m + (1 - m) * null_text, emb * mp + (1 - mp) * null_pos], -1)
    _check(out, ref, 1e-3, 2e-3, f"fourier_embed D{D} mode{mode} dropped={dropped}")


def test_plms_update_and_mean(cuda_device):
    ops = _ops()
    n = (4, 4, 64, 64)
    x, ec, eu, o1, o2, o3 = (_randn(n, cuda_device, 1.0, i) for i in range(6))
    a_t, a_prev = 0.31, 0.42
    s1 = math.sqrt(1 - a_t)
    e_out = torch.empty_like(x)
    x_out = torch.empty_like(x)
    ops.plms_update(x, ec, eu, 7.5, [o1, o2, o3], [55 / 24, -59 / 24, 37 / 24, -9 / 24], a_t, a_prev, s1, e_out, x_out)
    e = eu + 7.5 * (ec - eu)
    ep = (55 * e - 59 * o1 + 37 * o2 - 9 * o3) / 24
    pred = (x - s1 * ep) / math.sqrt(a_t)
    ref = math.sqrt(a_prev) * pred + math.sqrt(1 - a_prev) * ep
    _check(e_out, e, 1e-6, 1e-6, "plms e")
    _check(x_out, ref, 1e-5, 1e-5, "plms x_prev")
    m = torch.empty_like(x)
    ops.latent_mean([x, ec, eu], m)
    _check(m, (x + ec + eu) / 3, 1e-6, 1e-6, "latent mean")


# --------------------------------------------------------------------------------------------
# ConvNeXt mask-encoder pieces (csrc/convnext.cu) and the GELU epilogue
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(16384, 96, 48), (1024, 384, 96), (256, 768, 3072), (300, 192, 384)])
def test_gemm_gelu_epilogue_small_k_n(cuda_device, M, N, K):
    """pwconv1 (GELU fused) and the patchify GEMMs: N < one tile, K = 48 < one 64-wide k-block."""
    ops = _ops()
    a = _h(_randn((M, K), cuda_device, 1.0, 1))
    w = _h(_randn((N, K), cuda_device, 1.0 / math.sqrt(K), 2))
    bias = _randn((N,), cuda_device, 0.5, 3)
    ref = F.gelu(a.float() @ w.float().t() + bias)
    _check(ops.gemm(a, w, bias, gelu=True), ref, 2e-3, 2e-3, f"gemm+gelu {M}x{N}x{K}")
    _check(ops.gemm(a, w, bias), a.float() @ w.float().t() + bias, 2e-3, 2e-3, f"gemm {M}x{N}x{K}")


@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 96), (1, 128, 128, 96), (2, 8, 8, 768), (1, 5, 7, 192)])
def test_dwconv7x7(cuda_device, B, H, W, C):
    ops = _ops()
    x = _h(_randn((B, C, H, W), cuda_device, 1.0, 4))
    w = _randn((C, 1, 7, 7), cuda_device, 1.0 / 7, 5)
    b = _randn((C,), cuda_device, 0.1, 6)
    ref = F.conv2d(x.float(), w, b, padding=3, groups=C).permute(0, 2, 3, 1).reshape(B * H * W, C)
    x16 = x.permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous()
    out = ops.dwconv7x7(x16, w.reshape(C, 49).t().contiguous(), b, B, H, W)
    _check(out, ref, 2e-3, 2e-3, f"dwconv7x7 {B}x{H}x{W}x{C}")


@pytest.mark.parametrize("B,H,W,C,p", [(2, 16, 16, 3, 4), (1, 8, 8, 96, 2), (2, 4, 12, 384, 2)])
def test_patchify(cuda_device, B, H, W, C, p):
    ops = _ops()
    x = _h(_randn((B, H, W, C), cuda_device, 1.0, 7))
    out = ops.patchify(x.reshape(B * H * W, C), B, H, W, C, p)
    ref = x.view(B, H // p, p, W // p, p, C).permute(0, 1, 3, 2, 4, 5).reshape(B * (H // p) * (W // p), p * p * C)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("S", [512, 256, 100])
def test_segs_inconv_and_seg_tokens(cuda_device, S):
    ops = _ops()
    B, CI, R = 2, 30, 512
    segs = torch.zeros((B, CI, S, S), device=cuda_device)
    segs[0, 0, S // 8: S // 2, S // 4: S // 2] = 1.0
    segs[0, 3, : S // 3, : S // 5] = 1.0          # sample 1 stays all zero
    w = _randn((3, CI, 3, 3), cuda_device, 0.2, 8)
    b = _randn((3,), cuda_device, 0.1, 9)
    y, seg_sum = ops.segs_inconv(segs, w, b, R)
    rs = F.interpolate(segs, R, mode="nearest")
    ref = F.conv2d(rs, w, b, padding=1).permute(0, 2, 3, 1).reshape(B * R * R, 3)
    _check(y, ref, 1e-3, 1e-3, f"segs_inconv S={S}")
    assert torch.allclose(seg_sum, rs.sum(dim=(1, 2, 3)), rtol=1e-5), (seg_sum, rs.sum(dim=(1, 2, 3)))
    # expanded zero view (what the null / box-only inputs carry): strides are honoured
    zv = torch.zeros((B, CI, 1, 1), device=cuda_device).expand(B, CI, S, S)
    y0, s0 = ops.segs_inconv(zv, w, b, R)
    assert float(s0.abs().max()) == 0.0
    _check(y0, b.view(1, 3).expand(B * R * R, 3), 1e-3, 1e-3, "segs_inconv zero view")
    # token reinterpretation: reshape(B, -1, T).permute(0, 2, 1) of the NCHW feature map
    P, C, T = 256, 768, 64
    feat = _h(_randn((B, C, 16, 16), cuda_device, 1.0, 10))
    pos = _randn((T, C * P // T), cuda_device, 0.5, 11)
    null_pos = _h(_randn((T, C * P // T), cuda_device, 1.0, 12))
    nhwc = feat.permute(0, 2, 3, 1).reshape(B * P, C).contiguous()
    out = ops.seg_tokens(nhwc, null_pos, pos, seg_sum, B, P, T).view(B, T, -1)
    ref0 = feat.float().reshape(B, -1, T).permute(0, 2, 1)[0] + pos
    _check(out[0], ref0, 1e-3, 1e-3, "seg_tokens has-seg sample")
    assert torch.equal(out[1], null_pos)


# --------------------------------------------------------------------------------------------
# LayerNorm folded across two GEMMs (idiff_gemm_args.ln_*): producer row statistics + consumer epilogue
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,C,mean_shift", [(4096, 320, 0.0), (1024, 640, 0.5), (300, 1280, -1.0), (64, 1280, 3.0)])
def test_gemm_layernorm_fold(cuda_device, M, C, mean_shift):
    """x = residual + gate * (a W0^T + b0) written by a producer GEMM with want_stats; then
    LN(x) W1^T + b1 (plain) and GEGLU(LN(x) Wg^T + bg) through the folded consumers.  Reference: torch fp32
    LayerNorm / linear on the fp16 x the producer stored.  Also: the statistics themselves, and the one-slot
    row_stats entry point."""
    ops = _ops()
    a = _h(_randn((M, C), cuda_device, 1.0, 1))
    w0 = _h(_randn((C, C), cuda_device, 1.0 / math.sqrt(C), 2))
    b0 = _randn((C,), cuda_device, 0.5, 3) + mean_shift
    res = _h(_randn((M, C), cuda_device, 1.0, 4))
    x, st = ops.gemm(a, w0, b0, residual=res, gate=0.7, want_stats=True)
    xf = x.float()
    # statistics: sum over slots == row sums of the stored x (up to fp16 rounding of x: the epilogue sums fp32)
    got_sum = st.t[:, :, 0].sum(0)
    got_sq = st.t[:, :, 1].sum(0)
    _check(got_sum, xf.sum(1), 1e-3, 2e-2 * math.sqrt(C) / 16, f"ln stats sum M{M} C{C}")
    _check(got_sq, (xf * xf).sum(1), 2e-3, 1e-2, f"ln stats sumsq M{M} C{C}")
    st1 = ops.row_stats(x)
    _check(st1.t[0, :, 0], xf.sum(1), 1e-5, 1e-3, "row_stats sum")
    _check(st1.t[0, :, 1], (xf * xf).sum(1), 1e-5, 1e-3, "row_stats sumsq")

    gamma = 1.0 + 0.3 * _randn((C,), cuda_device, 1.0, 5)
    beta = 0.2 * _randn((C,), cuda_device, 1.0, 6)
    eps = 1e-5
    ln_ref = F.layer_norm(xf, (C,), gamma, beta, eps)
    # plain consumer (QKV-like, N = 3C, no bias)
    w1 = _randn((3 * C, C), cuda_device, 1.0 / math.sqrt(C), 7)
    f1 = ops.fold_layernorm(w1, None, gamma, beta, eps)
    ref1 = ln_ref @ w1.t()
    for name, s in (("producer stats", st), ("row_stats", st1)):
        out1 = ops.gemm(x, f1.w, f1.bias, ln=(s, f1.colsum, f1.eps))
        _check(out1, ref1, 3e-3, 4e-3, f"ln-fold plain M{M} C{C} ({name})")
    # GEGLU consumer (N = 8C, bias)
    from instancediffusion_b200.packing import pack_geglu
    wg = _randn((8 * C, C), cuda_device, 1.0 / math.sqrt(C), 8)
    bg = _randn((8 * C,), cuda_device, 0.3, 9)
    fg = ops.fold_layernorm(wg, bg, gamma, beta, eps, pack=pack_geglu)
    outg = ops.gemm(x, fg.w, fg.bias, geglu=True, ln=(st, fg.colsum, fg.eps))
    h = ln_ref @ wg.t() + bg
    v, g = h.chunk(2, dim=-1)
    _check(outg, v * F.gelu(g), 3e-3, 4e-3, f"ln-fold geglu M{M} C{C}")


def test_gemm_stats_from_long_k_producer(cuda_device):
    """FF out-projection at C=1280 (K = 5120 > 2560): the 8-warp direct epilogue also leaves row statistics."""
    ops = _ops()
    M, C, K = 2048, 1280, 5120
    a = _h(_randn((M, K), cuda_device, 1.0, 1))
    w = _h(_randn((C, K), cuda_device, 1.0 / math.sqrt(K), 2))
    b = _randn((C,), cuda_device, 0.5, 3)
    res = _h(_randn((M, C), cuda_device, 1.0, 4))
    x, st = ops.gemm(a, w, b, residual=res, want_stats=True)
    xf = x.float()
    _check(st.t[:, :, 0].sum(0), xf.sum(1), 1e-3, 0.1, "long-K stats sum")
    _check(st.t[:, :, 1].sum(0), (xf * xf).sum(1), 2e-3, 1e-2, "long-K stats sumsq")
    ref = res.float() + a.float() @ w.float().t() + b
    _check(x, ref, 2e-3, 2e-3, "long-K producer output")
