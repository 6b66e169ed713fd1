"""Native AutoencoderKL and CLIP text tower (csrc/vae.cu, csrc/clip.cu, through the C ABI) against their CPU fp32 oracles:
the oracle's restated AutoencoderKL (oracle/sd15.py) and the transformers CLIPTextModel the reference itself loads
(modules/models/__init__.py:135), on the same seeded random-init weights.  Reference call sites:
modules/inversion/diffusion_inversion.py:183-247."""
import pytest
import torch

pytestmark = pytest.mark.gpu

PROMPTS = ["a cat sitting next to a mirror", "a tiger sitting next to a mirror", "", "a photo of a very long sentence " * 12]


@pytest.fixture(scope="module")
def vae_weights():
    from eta_inversion_b200 import synthetic as syn
    return syn.random_state_dict(syn.vae_param_spec(), 1)


@pytest.fixture(scope="module")
def oracle_vae(vae_weights):
    from oracle import sd15
    v = sd15.AutoencoderKL().eval().requires_grad_(False)
    v.load_state_dict(vae_weights, strict=True)
    return v


@pytest.fixture(scope="module")
def vae_io(oracle_vae):
    """(image, oracle mean, latent, oracle decoded image) -- the oracle runs once per module (CPU fp32, ~10 s)."""
    from eta_inversion_b200 import synthetic as syn
    torch.set_num_threads(max(1, torch.get_num_threads()))
    img = syn.synthetic_image(0)
    with torch.no_grad():
        mean = oracle_vae.encode(img)["latent_dist"].mean
        g = torch.Generator().manual_seed(7)
        z = mean + 0.3 * torch.randn(mean.shape, generator=g)
        dec = oracle_vae.decode(z)["sample"]
    return img, mean, z, dec


def test_vae_fp32_matches_oracle(vae_weights, vae_io):
    from eta_inversion_b200.vae import VAEEngine
    img, mean, z, dec = vae_io
    vae = VAEEngine(vae_weights, dtype=torch.float32, device="cuda:0", max_batch=2)
    m = vae.encode(img.cuda())["latent_dist"].mean.cpu()
    err_e = (m - mean).abs().max().item()
    d = vae.decode(z.cuda())["sample"].cpu()
    err_d = (d - dec).abs().max().item()
    print(f"vae fp32: encode mean max-abs {err_e:.2e} (|mean| max {mean.abs().max():.2f}), decode max-abs {err_d:.2e} "
          f"(|image| max {dec.abs().max():.2f}); launches {vae.launch_count}")
    assert err_e < 1e-3 * max(1.0, mean.abs().max().item())
    assert err_d < 1e-3 * max(1.0, dec.abs().max().item())
    # batch of two different inputs == the two single calls (batch-invariant kernels), and a second call reproduces the first
    z2 = torch.cat([z, z.flip(-1)]).cuda()
    d2 = vae.decode(z2)["sample"]
    assert torch.equal(d2[0].cpu(), d[0])
    assert torch.equal(vae.decode(z2[1:])["sample"][0], d2[1])
    vae.close()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_vae_16bit_psnr(vae_weights, vae_io, dtype):
    from eta_inversion_b200.metrics import psnr
    from eta_inversion_b200.vae import VAEEngine
    img, mean, z, dec = vae_io
    vae = VAEEngine(vae_weights, dtype=dtype, device="cuda:0", max_batch=1)
    m = vae.encode(img.cuda().to(dtype))["latent_dist"].mean.float().cpu()
    rel = ((m - mean).norm() / mean.norm()).item()
    d = vae.decode(z.cuda().to(dtype))["sample"].float().cpu()
    p = psnr(d.clamp(-1, 1), dec.clamp(-1, 1))
    print(f"vae {dtype}: encode rel-L2 {rel:.2e}, decode PSNR {p:.1f} dB vs the fp32 oracle")
    assert rel < (2e-2 if dtype == torch.float16 else 6e-2)
    assert p >= 35.0
    vae.close()


def _ids(prompts):
    from eta_inversion_b200.models import SyntheticTokenizer
    return SyntheticTokenizer()(prompts).input_ids


def test_clip_fp32_matches_transformers():
    from eta_inversion_b200 import synthetic as syn
    from eta_inversion_b200.vae import CLIPTextEngine
    ref_model = syn.make_text_encoder(0)
    ids = _ids(PROMPTS)
    with torch.no_grad():
        ref = ref_model(ids)[0]
    clip = CLIPTextEngine.from_transformers(ref_model, dtype=torch.float32, device="cuda:0", max_batch=4)
    out = clip(ids)[0].cpu()
    err = (out - ref).abs().max().item()
    print(f"clip fp32: last_hidden_state max-abs {err:.2e} (|ref| max {ref.abs().max():.2f}); launches {clip.launch_count}")
    assert err < 2e-4 * max(1.0, ref.abs().max().item())
    # one prompt at a time (first call of a batch size is eager + captured, later ones replay the graph) == the batched pass
    for _ in range(2):
        for i in range(len(PROMPTS)):
            assert torch.equal(clip(ids[i:i + 1])[0][0].cpu(), out[i])
    clip.close()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_clip_16bit_close(dtype):
    from eta_inversion_b200 import synthetic as syn
    from eta_inversion_b200.vae import CLIPTextEngine
    ref_model = syn.make_text_encoder(0)
    ids = _ids(PROMPTS[:2])
    with torch.no_grad():
        ref = ref_model(ids)[0]
    clip = CLIPTextEngine.from_transformers(ref_model, dtype=dtype, device="cuda:0", max_batch=2)
    out = clip(ids)[0].float().cpu()
    rel = ((out - ref).norm() / ref.norm()).item()
    print(f"clip {dtype}: rel-L2 {rel:.2e}")
    assert rel < (1e-2 if dtype == torch.float16 else 4e-2)
    clip.close()
