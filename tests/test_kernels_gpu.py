"""GPU parity of every exported kernel (through the C ABI) against plain PyTorch fp32 on the same seeded inputs.
Tolerances: fp32 paths 1e-4 relative to the output scale; 16-bit tensor-core paths are compared against an fp32
reference computed from the SAME 16-bit-rounded inputs, tolerance = output rounding (2^-10 for f16, 2^-7 for bf16)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale)


def _relerr(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-12)).item()


TOL = {torch.float32: 2e-5, torch.float16: 3e-3, torch.bfloat16: 2e-2}


@pytest.mark.parametrize("M,N,K", [(4096, 320, 320), (154, 640, 768), (64, 1280, 2560), (300, 96, 200), (1, 4, 36)])
def test_gemm_simt_fp32(M, N, K):
    from eta_inversion_b200 import engine as E
    A, W, b, r = _rand((M, K), 1).cuda(), _rand((N, K), 2, K ** -0.5).cuda(), _rand((N,), 3).cuda(), _rand((M, N), 4).cuda()
    out = E.gemm(A, W, b, r, math_mode=E.MATH_SIMT)
    ref = (A.double() @ W.double().T + b.double() + r.double()).float()
    assert _relerr(out, ref) < TOL[torch.float32]


def test_gemm_simt_geglu():
    from eta_inversion_b200 import engine as E
    M, C = 512, 320
    A, W, b = _rand((M, C), 1).cuda(), _rand((8 * C, C), 2, C ** -0.5).cuda(), _rand((8 * C,), 3).cuda()
    # interleave (value, gate) rows as the engine packs them
    Wi = torch.stack([W[:4 * C], W[4 * C:]], 1).reshape(8 * C, C).contiguous()
    bi = torch.stack([b[:4 * C], b[4 * C:]], 1).reshape(8 * C).contiguous()
    out = E.gemm(A, Wi, bi, geglu=True, math_mode=E.MATH_SIMT)
    h = A @ W.T + b
    ref = h[:, :4 * C] * F.gelu(h[:, 4 * C:])
    assert _relerr(out, ref) < 5e-5


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(4096, 320, 320), (16384, 1280, 320), (154, 24960, 768), (64, 1280, 2560),
                                   (1000, 640, 640), (256, 128, 64), (128, 160, 5120)])
def test_gemm_tc(dtype, M, N, K):
    from eta_inversion_b200 import engine as E
    A, W = _rand((M, K), 1).to(dtype).cuda(), _rand((N, K), 2, K ** -0.5).to(dtype).cuda()
    b, r = _rand((N,), 3).to(dtype).cuda(), _rand((M, N), 4).to(dtype).cuda()
    out = E.gemm(A, W, b, r)
    ref = A.float() @ W.float().T + b.float() + r.float()
    assert _relerr(out.float(), ref) < TOL[dtype]
    out2 = E.gemm(A, W)
    assert _relerr(out2.float(), A.float() @ W.float().T) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float16])
def test_gemm_tc_geglu(dtype):
    from eta_inversion_b200 import engine as E
    M, C = 1024, 320
    A, W, b = _rand((M, C), 1).to(dtype).cuda(), _rand((8 * C, C), 2, C ** -0.5).to(dtype).cuda(), _rand((8 * C,), 3).to(dtype).cuda()
    Wi = torch.stack([W[:4 * C], W[4 * C:]], 1).reshape(8 * C, C).contiguous()
    bi = torch.stack([b[:4 * C], b[4 * C:]], 1).reshape(8 * C).contiguous()
    out = E.gemm(A, Wi, bi, geglu=True)
    h = A.float() @ W.float().T + b.float()
    ref = h[:, :4 * C] * F.gelu(h[:, 4 * C:])
    assert _relerr(out.float(), ref) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(20077, 320, 640), (9900, 960, 320), (37 * 256 + 1, 320, 64), (16384, 1280, 1280)])
def test_gemm_tc_large(dtype, M, N, K):
    """Many tiles per persistent CTA (several rounds of the double-buffered accumulator), ragged M tails."""
    from eta_inversion_b200 import engine as E
    A, W = _rand((M, K), 1).to(dtype).cuda(), _rand((N, K), 2, K ** -0.5).to(dtype).cuda()
    b, r = _rand((N,), 3).to(dtype).cuda(), _rand((M, N), 4).to(dtype).cuda()
    out = E.gemm(A, W, b, r)
    ref = A.float() @ W.float().T + b.float() + r.float()
    assert _relerr(out.float(), ref) < TOL[dtype]
    # per-row check on the first / last rows of tiles
    for m in (0, 127, 128, 255, M - 1, M - 130 if M > 130 else 0):
        assert _relerr(out[m].float(), ref[m]) < 4 * TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(65536, 320, 320), (32768 + 77, 960, 320), (16384, 640, 2560), (65536, 320, 1280),
                                   (4096, 3840, 1280)])
def test_gemm_tc_wide_tiles(dtype, M, N, K):
    """Shapes the host routes to the 128 x 320 tile (two UMMAs of N = 160 per k-step, single-buffered accumulators)."""
    from eta_inversion_b200 import engine as E
    A, W = _rand((M, K), 1).to(dtype).cuda(), _rand((N, K), 2, K ** -0.5).to(dtype).cuda()
    b, r = _rand((N,), 3).to(dtype).cuda(), _rand((M, N), 4).to(dtype).cuda()
    out = E.gemm(A, W, b, r)
    ref = torch.addmm(r.float(), A.float(), W.float().T) + b.float()
    assert _relerr(out.float(), ref) < TOL[dtype]
    for m in (0, 127, 128, M - 1):
        assert _relerr(out[m].float(), ref[m]) < 4 * TOL[dtype]
    for n0 in (0, 160, N - 160):  # every accumulator half of the first / last tile column
        assert _relerr(out[:, n0:n0 + 160].float(), ref[:, n0:n0 + 160]) < 2 * TOL[dtype]


@pytest.mark.parametrize("M,C", [(8192, 320), (65536, 320), (16384, 640)])
def test_gemm_tc_large_geglu(M, C):
    from eta_inversion_b200 import engine as E
    dtype = torch.float16
    A, W, b = _rand((M, C), 1).to(dtype).cuda(), _rand((8 * C, C), 2, C ** -0.5).to(dtype).cuda(), _rand((8 * C,), 3).to(dtype).cuda()
    Wi = torch.stack([W[:4 * C], W[4 * C:]], 1).reshape(8 * C, C).contiguous()
    bi = torch.stack([b[:4 * C], b[4 * C:]], 1).reshape(8 * C).contiguous()
    out = E.gemm(A, Wi, bi, geglu=True)
    h = A.float() @ W.float().T + b.float()
    ref = h[:, :4 * C] * F.gelu(h[:, 4 * C:])
    del h
    assert _relerr(out.float(), ref) < TOL[dtype]


def _conv_ref(x_nhwc, w_oihw, bias, stride):
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2).float(), w_oihw.float(), bias.float(), stride=stride, padding=1)
    return y.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("B,H,Ci,Co,stride", [(2, 64, 4, 320, 1), (1, 32, 320, 4, 1), (2, 16, 64, 96, 1), (3, 16, 64, 64, 2),
                                              (1, 8, 128, 64, 1)])
def test_conv3x3_simt_fp32(B, H, Ci, Co, stride):
    from eta_inversion_b200 import engine as E
    x, w, b = _rand((B, H, H, Ci), 1).cuda(), _rand((Co, Ci, 3, 3), 2, (9 * Ci) ** -0.5).cuda(), _rand((Co,), 3).cuda()
    wp = w.permute(0, 2, 3, 1).contiguous()
    out = E.conv3x3(x, wp, b, stride=stride, math_mode=E.MATH_SIMT)
    assert _relerr(out, _conv_ref(x, w, b, stride)) < TOL[torch.float32]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,H,Ci,Co,stride", [(2, 64, 320, 320, 1), (1, 32, 640, 640, 1), (3, 16, 1280, 640, 1),
                                              (4, 8, 2560, 1280, 1), (1, 8, 1280, 1280, 1), (2, 64, 320, 320, 2),
                                              (3, 16, 1280, 1280, 2), (2, 32, 960, 640, 1)])
def test_conv3x3_tc(dtype, B, H, Ci, Co, stride):
    from eta_inversion_b200 import engine as E
    x = _rand((B, H, H, Ci), 1).to(dtype).cuda()
    w = _rand((Co, Ci, 3, 3), 2, (9 * Ci) ** -0.5).to(dtype).cuda()
    b = _rand((Co,), 3).to(dtype).cuda()
    wp = w.permute(0, 2, 3, 1).contiguous()
    out = E.conv3x3(x, wp, b, stride=stride)
    ref = _conv_ref(x, w, b, stride)
    assert _relerr(out.float(), ref) < TOL[dtype]
    res = _rand(tuple(ref.shape), 5).to(dtype).cuda()
    out = E.conv3x3(x, wp, b, residual=res, stride=stride)
    assert _relerr(out.float(), ref + res.float()) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,H,Ci,Co", [(5, 64, 320, 320), (9, 32, 640, 640), (12, 16, 1280, 1280), (40, 8, 1280, 1280),
                                        (11, 16, 2560, 1280), (16, 64, 320, 320), (16, 64, 960, 320), (16, 32, 1280, 640),
                                        (16, 16, 2560, 1280), (16, 8, 2560, 1280)])
def test_conv3x3_tc_large(dtype, B, H, Ci, Co):
    """Implicit-GEMM conv at co-batched sizes (odd image counts: the last 128-pixel box is partly out of range)."""
    from eta_inversion_b200 import engine as E
    x = _rand((B, H, H, Ci), 1).to(dtype).cuda()
    w = _rand((Co, Ci, 3, 3), 2, (9 * Ci) ** -0.5).to(dtype).cuda()
    b = _rand((Co,), 3).to(dtype).cuda()
    wp = w.permute(0, 2, 3, 1).contiguous()
    res = _rand((B, H, H, Co), 5).to(dtype).cuda()
    out = E.conv3x3(x, wp, b, residual=res)
    ref = _conv_ref(x, w, b, 1) + res.float()
    assert _relerr(out.float(), ref) < TOL[dtype]
    for img in (0, B - 1):
        assert _relerr(out[img].float(), ref[img]) < 4 * TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,HW,C,silu", [(2, 4096, 320, True), (1, 1024, 960, True), (3, 64, 2560, True), (2, 256, 1920, False),
                                         (1, 4096, 640, False),
                                         (2, 4096, 960, True),     # fp32: slab not resident (apply re-reads x); 16-bit: 1 CTA/SM
                                         (2, 1024, 1280, True), (5, 256, 640, False), (3, 1024, 1920, True),
                                         (2, 100, 320, True),      # rows not divisible by the cluster size
                                         (1, 16384, 128, True)])   # VAE-sized: two-launch path
def test_groupnorm(dtype, B, HW, C, silu):
    from eta_inversion_b200 import engine as E
    x = (_rand((B, HW, C), 1) * 2 + 0.5).to(dtype).cuda()
    g, b = (1 + 0.1 * _rand((C,), 2)).to(dtype).cuda(), (0.1 * _rand((C,), 3)).to(dtype).cuda()
    out = E.groupnorm(x, g, b, 32, 1e-5, silu)
    ref = F.group_norm(x.float().permute(0, 2, 1), 32, g.float(), b.float(), 1e-5).permute(0, 2, 1)
    if silu:
        ref = F.silu(ref)
    assert _relerr(out.float(), ref) < {torch.float32: 2e-5, torch.float16: 2e-3, torch.bfloat16: 1.6e-2}[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_groupnorm_is_batch_invariant_and_deterministic(dtype):
    """A row's result must not depend on what it is batched with (lock-step groups) nor vary run to run (A/B/A)."""
    from eta_inversion_b200 import engine as E
    HW, C = 1024, 640
    x = (_rand((5, HW, C), 7) * 2 + 0.5).to(dtype).cuda()
    g, b = (1 + 0.1 * _rand((C,), 2)).to(dtype).cuda(), (0.1 * _rand((C,), 3)).to(dtype).cuda()
    full = E.groupnorm(x, g, b, 32, 1e-5, True)
    assert torch.equal(full, E.groupnorm(x, g, b, 32, 1e-5, True))
    for r in (0, 3):
        assert torch.equal(full[r:r + 1], E.groupnorm(x[r:r + 1].contiguous(), g, b, 32, 1e-5, True))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("M,C", [(4096, 320), (1000, 640), (77, 1280)])
def test_layernorm(dtype, M, C):
    from eta_inversion_b200 import engine as E
    x = (_rand((M, C), 1) * 3 - 1).to(dtype).cuda()
    g, b = (1 + 0.1 * _rand((C,), 2)).to(dtype).cuda(), (0.1 * _rand((C,), 3)).to(dtype).cuda()
    out = E.layernorm(x, g, b, 1e-5)
    ref = F.layer_norm(x.float(), (C,), g.float(), b.float(), 1e-5)
    assert _relerr(out.float(), ref) < (2e-5 if dtype == torch.float32 else 2e-3)


@pytest.mark.parametrize("N,d", [(4096, 40), (1024, 80), (256, 160), (64, 160), (200, 40)])
def test_attention_simt_fp32(N, d):
    from eta_inversion_b200 import engine as E
    B, heads = 3, 8
    qkv = _rand((B, N, 3 * heads * d), 1).cuda()
    C = heads * d
    q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    rows = ([0, 0, 2], [0, 0, 1], [0, 1, 2])  # exercises the (q,k,v) source remap
    out = E.attention(q, k, v, heads, rows=rows, math_mode=E.MATH_SIMT)

    def sh(t):
        return t.reshape(B, N, heads, d).permute(0, 2, 1, 3)
    qq, kk, vv = sh(q)[list(rows[0])], sh(k)[list(rows[1])], sh(v)[list(rows[2])]
    ref = F.scaled_dot_product_attention(qq.double(), kk.double(), vv.double()).permute(0, 2, 1, 3).reshape(B, N, C)
    assert _relerr(out, ref.float()) < 2e-5


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N,d", [(4096, 40), (1024, 80), (256, 160), (64, 160), (200, 40), (1024, 40)])
def test_attention_tc(dtype, N, d):
    """tcgen05 flash attention (TMA zero-padded head dims, MN-major V operand) vs fp32 SDPA on the same inputs."""
    from eta_inversion_b200 import engine as E
    B, heads = 3, 8
    C = heads * d
    qkv = _rand((B, N, 3 * C), 1).to(dtype).cuda()
    q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    rows = ([0, 0, 2], [0, 0, 1], [0, 1, 2])
    out = E.attention(q, k, v, heads, rows=rows)

    def sh(t):
        return t.float().reshape(B, N, heads, d).permute(0, 2, 1, 3)
    qq, kk, vv = sh(q)[list(rows[0])], sh(k)[list(rows[1])], sh(v)[list(rows[2])]
    ref = F.scaled_dot_product_attention(qq, kk, vv).permute(0, 2, 1, 3).reshape(B, N, C)
    err = _relerr(out.float(), ref)
    print(f"attention_tc {dtype} N={N} d={d}: rel err {err:.2e}")
    assert err < (4e-3 if dtype == torch.float16 else 2e-2)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N", [4096, 1000])
def test_attention_tc_growing_logits(dtype, N):
    """Keys whose magnitude ramps up along the sequence: the running row maximum keeps growing by more than the lazy
    rescale threshold (2^8), so the in-TMEM rescale of the d = 40 kernel runs many times per row, on some rows only."""
    from eta_inversion_b200 import engine as E
    B, heads, d = 2, 8, 40
    C = heads * d
    g = torch.Generator().manual_seed(7)
    q = torch.randn((B, N, C), generator=g)
    k = torch.randn((B, N, C), generator=g) * torch.linspace(0.2, 14.0, N).reshape(1, N, 1)
    k[:, :, :C // 2] *= 0.1  # half of the heads stay flat: no rescale there
    v = torch.randn((B, N, C), generator=g)
    q, k, v = (t.to(dtype).cuda() for t in (q, k, v))
    out = E.attention(q, k, v, heads)

    def sh(t):
        return t.float().reshape(B, N, heads, d).permute(0, 2, 1, 3)
    ref = F.scaled_dot_product_attention(sh(q), sh(k), sh(v)).permute(0, 2, 1, 3).reshape(B, N, C)
    err = _relerr(out.float(), ref)
    print(f"attention_tc growing {dtype} N={N}: rel err {err:.2e}")
    assert err < (4e-3 if dtype == torch.float16 else 2e-2)


def test_scheduler_step_matches_closed_form():
    from eta_inversion_b200 import engine as E
    n, Esz = 2, 4 * 64 * 64
    eps, x = _rand((2 * n, 4, 64, 64), 1).cuda(), _rand((n, 4, 64, 64), 2).cuda()
    a_t, a_p, g, eta = 0.3, 0.45, 7.5, 0.37
    var = ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)
    mask = (_rand((1, 4, 64, 64), 3) > 0).float().cuda()
    cand = _rand((10, 1, 4, 64, 64), 4).cuda()
    xinv = _rand((1, 4, 64, 64), 5).cuda()
    losses, best = E.eta_noise_losses(eps, x, xinv, a_t, a_p, g, eta, var, cand)
    e = eps[:n] + g * (eps[n:] - eps[:n])
    x0 = (x - math.sqrt(1 - a_t) * e) / math.sqrt(a_t)
    sig = eta * math.sqrt(var)
    rec = math.sqrt(a_p) * x0[:1] + math.sqrt(1 - a_p - sig ** 2) * e[:1]
    zopt = (xinv - rec) / sig
    ref_losses = (cand - zopt).square().reshape(10, -1).mean(1)
    assert torch.allclose(losses, ref_losses, rtol=1e-4)
    assert int(best.item()) == int(ref_losses.argmin().item())
    out, e_out = E.cfg_ddim_step(eps, x, a_t, a_p, g, eta, var, eta_map=mask, noise_cand=cand, losses=losses, pin_src=xinv,
                                 want_eps=True)
    sigm = eta * mask * math.sqrt(var)
    ref = math.sqrt(a_p) * x0 + (1 - a_p - sigm ** 2).sqrt() * e + sigm * cand[ref_losses.argmin()]
    ref[:1] = xinv
    assert torch.allclose(e_out, e, atol=1e-5)
    assert _relerr(out, ref) < 2e-6
    # inversion direction, no cfg, eta = 0: the DDIM inverse of scheduling_ddim_inverse.py:94-98
    out = E.cfg_ddim_step(eps[:n], x, a_p, a_t)
    x0 = (x - math.sqrt(1 - a_p) * eps[:n]) / math.sqrt(a_p)
    ref = math.sqrt(a_t) * x0 + math.sqrt(1 - a_t) * eps[:n]
    assert (out - ref).abs().max().item() < 1e-5
    # forward then backward with the same eps is the identity at eta = 0
    back = E.cfg_ddim_step(eps[:n], out, a_t, a_p)
    assert (back - x).abs().max().item() < 1e-4


@pytest.mark.parametrize("shape,q,l1,dup", [((2, 4, 64, 64), 0.7, False, False), ((1, 4, 64, 64), 0.7, True, False),
                                           ((2, 4, 64, 64), 0.5, False, True), ((1, 1000), 0.999, True, False),
                                           ((3, 4, 64, 64), 0.0001, False, False), ((1, 4, 64, 64), 1.0, False, False),
                                           ((2, 4, 64, 64), -0.05, True, False), ((1, 7), 0.3, False, True)])
def test_prox_guidance_matches_torch_quantile(shape, q, l1, dup):
    """etai_prox_guidance (radix select + soft threshold + guidance in one launch) against the reference's torch lines
    (proximal_negative_prompt_inversion.py:84-113): same threshold as torch.quantile -- the same order statistics and the
    same float32 rank arithmetic -- and the same thresholded CFG output."""
    from eta_inversion_b200 import engine as E
    u, c = _rand(shape, 11).cuda(), _rand(shape, 12).cuda()
    if dup:  # many equal |delta| values: the successor search must step over the run of duplicates
        u, c = (u * 4).round() / 4, (c * 4).round() / 4
    g = 7.5
    out, thr = E.prox_guidance(u, c, g, q, l1=l1, want_thr=True)
    delta = c - u
    thr_ref = delta.abs().quantile(q) if q > 0 else torch.tensor(-q, device="cuda")
    assert torch.allclose(thr[0], thr_ref.float(), rtol=1e-6, atol=0), (thr.item(), thr_ref.item())
    t = thr[0]  # the thresholding itself is compared at the kernel's own threshold (a 1-ulp lerp difference moves no element)
    d = delta - delta.clamp(-t, t)
    if l1:
        d = torch.where(d > 0, d - t, d)
        d = torch.where(d < 0, d + t, d)
    ref = u + g * d
    assert torch.equal(out, ref)
    if q > 0 and not dup:  # the fraction of surviving elements is what the quantile promises
        frac = (d != 0).float().mean().item()
        assert abs(frac - (1 - q)) <= 2.0 / delta.numel() + (0.02 if l1 else 0.0)


def test_prox_guidance_nan_propagates():
    from eta_inversion_b200 import engine as E
    u, c = _rand((1, 4, 64, 64), 1).cuda(), _rand((1, 4, 64, 64), 2).cuda()
    c[0, 1, 2, 3] = float("nan")
    out, thr = E.prox_guidance(u, c, 7.5, 0.7, want_thr=True)
    assert torch.isnan(thr).all() and torch.isnan(out).all()  # torch.quantile returns NaN, clamp(NaN bounds) spreads it


def test_ddim_step_stochastic_rows_get_independent_noise():
    """DDIMScheduler.step(eta > 0, variance_noise=None) with B > 1: diffusers draws randn of the SAMPLE's shape, so every row
    has its own noise; an explicit per-row variance_noise [B,C,H,W] is honoured row by row."""
    from eta_inversion_b200.models import sd_scheduler
    sch = sd_scheduler()
    sch.set_timesteps(10)
    t = sch.timesteps[3]
    x, eps = _rand((2, 4, 64, 64), 7).cuda(), _rand((2, 4, 64, 64), 8).cuda()
    eps[1], x[1] = eps[0], x[0]  # identical rows: any difference in the output is the noise
    g = torch.Generator(device="cuda").manual_seed(0)
    out = sch.step(eps, t, x, eta=1.0, generator=g).prev_sample
    assert (out[0] - out[1]).abs().max().item() > 1e-2
    noise = _rand((2, 4, 64, 64), 9).cuda()
    o2 = sch.step(eps, t, x, eta=1.0, variance_noise=noise).prev_sample
    a_t, a_p = sch.alpha(int(t)), sch.alpha(sch.prev_timestep(t))
    sig = float(sch._get_variance(int(t), sch.prev_timestep(t))) ** 0.5
    x0 = (x - math.sqrt(1 - a_t) * eps) / math.sqrt(a_t)
    ref = math.sqrt(a_p) * x0 + math.sqrt(max(1 - a_p - sig ** 2, 0.0)) * eps + sig * noise
    assert (o2 - ref).abs().max().item() < 1e-4


def test_errors_are_loud():
    from eta_inversion_b200 import engine as E
    with pytest.raises(RuntimeError):
        E.gemm(torch.zeros(4, 8), torch.zeros(4, 8))  # CPU tensors: no fallback
    A = torch.zeros(8, 24, device="cuda", dtype=torch.float16)
    with pytest.raises(RuntimeError, match="tcgen05"):
        E.gemm(A, torch.zeros(8, 24, device="cuda", dtype=torch.float16))  # K not a multiple of 64


# ---------------------------------------------------------------------------------------------------------------------
# control-aware cross-attention (etai_cross_attention): the fused softmax -> prompt-to-prompt edit -> store -> PV of ONE
# layer against the explicit computation of the reference (ptp_utils.py:238-253 + ptp.py:205-274), heads materialised.
# ---------------------------------------------------------------------------------------------------------------------
def _ref_cross_attention(q, k, v, heads, scale, pairs, mapper, blend_a, equalizer, alpha, store_rows):
    B, N, C = q.shape
    L, d = k.shape[1], C // heads
    def split(t):
        return t.float().reshape(t.shape[0], t.shape[1], heads, d).permute(0, 2, 1, 3)  # [B,h,n,d]
    P = torch.softmax(split(q) @ split(k).transpose(-1, -2) * scale, dim=-1)            # [B,h,N,L]
    for p, (br, tr) in enumerate(pairs or []):
        base, tgt = P[br], P[tr]
        rep = base @ mapper[p].float()                                                   # einsum('hpw,wn->hpn')
        f = equalizer[p].float() * (blend_a[p].float() * rep + (1 - blend_a[p].float()) * tgt)
        P[tr] = alpha[p].float() * f + (1 - alpha[p].float()) * tgt                      # no renormalisation
    out = (P @ split(v)).permute(0, 2, 1, 3).reshape(B, N, C)
    store = None if store_rows is None else torch.stack([P[r].sum(0) for r in store_rows])  # post-edit, summed over heads
    return out, store


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N,d,mode", [(256, 160, "replace"), (256, 160, "refine"), (1024, 80, "replace"), (4096, 40, "refine"),
                                      (64, 160, "replace"), (256, 160, "none")])
def test_cross_attention_control(dtype, N, d, mode):
    from eta_inversion_b200 import engine as E
    B, heads, L, X = 6, 8, 77, 768
    C = heads * d
    q = _rand((B, N, C), 1).to(dtype).cuda()
    ld = 2 * C + 64                                     # K at column 32, V at column 32 + C of a wider buffer
    kv = _rand((B, L, ld), 2).to(dtype).cuda()
    koff, voff = 32, 32 + C
    k, v = kv[:, :, koff:koff + C], kv[:, :, voff:voff + C]
    scale = 2.5 * d ** -0.5                             # peaky rows: the edit must act on structured probabilities
    pairs = None if mode == "none" else [(2, 3), (4, 5)]
    gen = torch.Generator().manual_seed(5)
    if mode == "replace":   # word-swap mapper: permutation-like with a few fractional rows (seq_aligner.get_replacement_mapper)
        mapper = torch.stack([torch.eye(L)[torch.randperm(L, generator=gen)] for _ in range(2)])
        mapper[:, 5, :] = 0
        mapper[:, 5, 5:8] = 1 / 3
        blend_a = torch.ones((2, L))
    elif mode == "refine":  # gather by index + per-token blend (ptp.py:247-259): one-hot mapper, alphas in {0,1}
        idx = torch.stack([torch.randperm(L, generator=gen) for _ in range(2)])
        mapper = torch.zeros((2, L, L))
        for p in range(2):
            mapper[p, idx[p], torch.arange(L)] = 1.0
        blend_a = (torch.rand((2, L), generator=gen) > 0.3).float()
    else:
        mapper = blend_a = None
    if pairs:
        equalizer = torch.ones((2, L))
        equalizer[:, 2] = 2.0                           # reweight one word (ptp.py:262-274)
        alpha = (torch.rand((2, L), generator=gen) > 0.25).float()
        mapper, blend_a, equalizer, alpha = [t.cuda().contiguous() for t in (mapper, blend_a, equalizer, alpha)]
    else:
        equalizer = alpha = None
    store_rows = [3, 2, 5, 0] if N <= 1024 else None
    acc = torch.full((len(store_rows), N, L), 0.5, device="cuda") if store_rows else None  # accumulated in place
    out = E.cross_attention(q, kv, heads, koff, voff, scale, pairs, mapper, blend_a, equalizer, alpha, store_rows, acc)
    ref, ref_store = _ref_cross_attention(q, k, v, heads, scale, pairs, mapper, blend_a, equalizer, alpha, store_rows)
    assert _relerr(out.float(), ref) < TOL[dtype]
    if store_rows:
        assert _relerr(acc - 0.5, ref_store) < TOL[dtype]
    # the SIMT fp32-math path on the same 16-bit inputs agrees as well
    if dtype != torch.float32:
        acc2 = torch.zeros_like(acc) if store_rows else None
        out2 = E.cross_attention(q, kv, heads, koff, voff, scale, pairs, mapper, blend_a, equalizer, alpha, store_rows, acc2,
                                 math_mode=E.MATH_SIMT)
        assert _relerr(out2.float(), ref) < TOL[dtype]
