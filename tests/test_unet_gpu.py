"""UNet forward parity on the GPU: engine (C ABI) vs the committed golden produced by the oracle UNet on CPU fp32
(oracle/run_reference.py::run_unet_fwd, same seeded weights and inputs)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _inputs():
    g = torch.Generator().manual_seed(1234)
    x = torch.randn((4, 4, 64, 64), generator=g)
    ctx = torch.randn((4, 77, 768), generator=g)
    return x, ctx


@pytest.mark.parametrize("t", [981, 1])
def test_unet_fp32_matches_golden(engine_fp32, t):
    x, ctx = _inputs()
    gold = torch.from_numpy(np.load(GOLDEN / "unet_fwd.npz")[f"eps_t{t}"])
    out = engine_fp32(x.cuda(), torch.tensor(t), encoder_hidden_states=ctx.cuda())["sample"].cpu()
    err = (out - gold).abs().max().item()
    print(f"fp32 UNet t={t}: max-abs err {err:.3e} (eps abs-mean {gold.abs().mean():.3f})")
    assert err < 5e-4  # north_star: latents within 1e-3 per step in fp32


def test_unet_fp32_batch_rows_independent(engine_fp32):
    x, ctx = _inputs()
    full = engine_fp32(x.cuda(), 501, encoder_hidden_states=ctx.cuda())["sample"]
    part = engine_fp32(x[1:3].cuda().contiguous(), 501, encoder_hidden_states=ctx[1:3].cuda().contiguous())["sample"]
    assert (full[1:3] - part).abs().max().item() < 1e-5
    again = engine_fp32(x.cuda(), 501, encoder_hidden_states=ctx.cuda())["sample"]
    assert torch.equal(full, again)  # bit-exact run to run (no atomics)


@pytest.mark.parametrize("t", [981, 1])
def test_unet_fp16_close_to_golden(engine_fp16, t):
    x, ctx = _inputs()
    gold = torch.from_numpy(np.load(GOLDEN / "unet_fwd.npz")[f"eps_t{t}"])
    out = engine_fp16(x.cuda(), torch.tensor(t), encoder_hidden_states=ctx.cuda())["sample"].cpu()
    rel = ((out - gold).norm() / gold.norm()).item()
    print(f"fp16 UNet t={t}: rel-L2 err {rel:.3e}, max-abs {(out - gold).abs().max():.3e}")
    assert rel < 2e-2
