"""Size-independent properties at the full sizes of the hot path (BASELINE configs: 64x64 latents, B up to 16) and the
error behaviour of the C ABI on a GPU box.  These complement the oracle comparisons, which run at sizes the CPU oracle
finishes in seconds."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0, dtype=torch.float16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).cuda()


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-12)).item()


@pytest.mark.parametrize("N,d", [(4096, 40), (1024, 80), (256, 160)])
def test_attention_is_invariant_to_key_order(N, d):
    """softmax(QK^T)V does not depend on the order of the keys: permuting (K, V) rows changes the tiling, the running
    maxima and every rescale decision of the flash kernels, but not the result."""
    from eta_inversion_b200 import engine as E
    B, heads = 8, 8
    C = heads * d
    q, k, v = _rand((B, N, C), 1), _rand((B, N, C), 2, 1.5), _rand((B, N, C), 3)
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(4)).cuda()
    a = E.attention(q, k, v, heads)
    b = E.attention(q, k[:, perm].contiguous(), v[:, perm].contiguous(), heads)
    assert _rel(a.float(), b.float()) < 4e-3


def test_attention_is_linear_in_v_at_full_size():
    from eta_inversion_b200 import engine as E
    B, heads, N, d = 16, 8, 4096, 40
    C = heads * d
    q, k = _rand((B, N, C), 1), _rand((B, N, C), 2)
    v1, v2 = _rand((B, N, C), 3), _rand((B, N, C), 4)
    lhs = E.attention(q, k, (0.5 * v1 + 0.25 * v2).contiguous(), heads).float()
    rhs = 0.5 * E.attention(q, k, v1, heads).float() + 0.25 * E.attention(q, k, v2, heads).float()
    assert _rel(lhs, rhs) < 6e-3


def test_attention_rows_of_a_constant_v_return_that_constant():
    """Every softmax row sums to one: with V = const the output is that constant whatever Q and K are (checks the
    row-sum "ones column", the lazy rescale and the masked tail of a ragged key count)."""
    from eta_inversion_b200 import engine as E
    B, heads, d = 4, 8, 40
    C = heads * d
    for N in (4096, 777):
        q, k = _rand((B, N, C), 1, 2.0), _rand((B, N, C), 2, 2.0)
        v = torch.full((B, N, C), 0.75, dtype=torch.float16, device="cuda")
        out = E.attention(q, k, v, heads).float()
        assert (out - 0.75).abs().max().item() < 2e-3


def test_conv_and_gemm_are_linear_at_full_size():
    from eta_inversion_b200 import engine as E
    B, H, Ci, Co = 16, 64, 320, 320
    x1, x2 = _rand((B, H, H, Ci), 1), _rand((B, H, H, Ci), 2)
    w = _rand((Co, 3, 3, Ci), 3, (9 * Ci) ** -0.5)
    zero = torch.zeros((Co,), dtype=torch.float16, device="cuda")
    lhs = E.conv3x3((x1 + x2).contiguous(), w, zero).float()
    rhs = E.conv3x3(x1, w, zero).float() + E.conv3x3(x2, w, zero).float()
    assert _rel(lhs, rhs) < 4e-3
    A1, A2, W = _rand((B * H * H, 320), 4), _rand((B * H * H, 320), 5), _rand((960, 320), 6, 320 ** -0.5)
    lhs = E.gemm((A1 - A2).contiguous(), W).float()
    rhs = E.gemm(A1, W).float() - E.gemm(A2, W).float()
    assert _rel(lhs, rhs) < 4e-3


def test_groupnorm_output_statistics_at_full_size():
    """Normalised groups have zero mean and unit variance (gamma = 1, beta = 0), at B = 16 where the partition into
    chunks spans several waves of CTAs."""
    from eta_inversion_b200 import engine as E
    B, HW, C, G = 16, 4096, 320, 32
    x = (_rand((B, HW, C), 1, 3.0, torch.float32) + 1.5).to(torch.float16)
    y = E.groupnorm(x, torch.ones(C, dtype=torch.float16, device="cuda"), torch.zeros(C, dtype=torch.float16, device="cuda"),
                    G, 1e-5, False).float().reshape(B, HW, G, C // G)
    assert y.mean(dim=(1, 3)).abs().max().item() < 2e-3
    assert (y.var(dim=(1, 3), unbiased=False) - 1).abs().max().item() < 5e-3


def test_ddim_inversion_round_trip_is_the_identity_for_a_fixed_eps():
    from eta_inversion_b200 import engine as E
    x, eps = _rand((16, 4, 64, 64), 1, dtype=torch.float32), _rand((16, 4, 64, 64), 2, dtype=torch.float32)
    a_lo, a_hi = 0.9, 0.35
    fwd = E.cfg_ddim_step(eps, x, a_lo, a_hi)      # z_t -> z_{t+1}
    back = E.cfg_ddim_step(eps, fwd, a_hi, a_lo)   # and back with the same eps
    assert (back - x).abs().max().item() < 2e-6


# ---- error behaviour at the ABI: negative code + message, never a crash or a silent fallback -------------------------
def test_errors_are_reported_not_swallowed():
    import eta_inversion_b200 as etai
    from eta_inversion_b200 import engine as E
    with pytest.raises(RuntimeError, match="unsupported dtype"):
        E.gemm(torch.zeros((8, 64), dtype=torch.int32, device="cuda"), torch.zeros((32, 64), dtype=torch.int32, device="cuda"))
    with pytest.raises(RuntimeError, match="etai"):
        E.groupnorm(_rand((1, 64, 100), 1), _rand((100,), 2), _rand((100,), 3), 32)        # C % 8, C % groups
    with pytest.raises(RuntimeError, match="etai"):
        E.attention(_rand((1, 64, 8 * 24), 1), _rand((1, 64, 8 * 24), 2), _rand((1, 64, 8 * 24), 3), 8,
                    math_mode=E.MATH_AUTO, rows=([5], [0], [0]))                               # source row out of range
    pipe, _ = etai.load_diffusion_model("synthetic-sd15", "cuda:0", variant="fp16", max_batch=2)
    lat = torch.zeros((3, 4, 64, 64), device="cuda")
    ctx = torch.zeros((3, 77, 768), device="cuda")
    with pytest.raises(RuntimeError, match="etai error"):
        pipe.unet(lat, 1, encoder_hidden_states=ctx)                                            # B > max_batch
    with pytest.raises(RuntimeError, match="etai"):
        pipe.unet(lat[:2], 1, encoder_hidden_states=torch.zeros((2, 77, 512), device="cuda"))  # wrong context width
    out = pipe.unet(lat[:2], 1, encoder_hidden_states=ctx[:2])["sample"]                        # the handle still works
    assert out.shape == (2, 4, 64, 64) and torch.isfinite(out).all()
