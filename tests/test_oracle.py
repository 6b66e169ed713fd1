"""Pins the oracle (CPU): (1) the restated diffusers arithmetic reproduces its committed golden on this machine,
(2) the self-contained loop restatement (oracle/ref_loop.py, which is what runs on the GPU box) reproduces the golden
trajectory that the reference's own UNMODIFIED modules/** produced (oracle/run_reference.py), (3) closed-form
self-checks of the scheduler algebra."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN


@pytest.fixture(scope="module")
def pipe(unet_weights):
    from eta_inversion_b200 import synthetic as syn
    from oracle import sd15
    return sd15.build_pipeline(unet_weights, syn.random_state_dict(syn.vae_param_spec(), 1), seed=0)


def test_oracle_unet_reproduces_golden(pipe):
    g = torch.Generator().manual_seed(1234)
    x, ctx = torch.randn((4, 4, 64, 64), generator=g), torch.randn((4, 77, 768), generator=g)
    gold = np.load(GOLDEN / "unet_fwd.npz")
    with torch.no_grad():
        out = pipe.unet(x[:2], torch.tensor(981), encoder_hidden_states=ctx[:2])["sample"]
    assert (out - torch.from_numpy(gold["eps_t981"])[:2]).abs().max().item() < 2e-5  # other CPU / thread count


def test_oracle_attention_paths_agree(pipe):
    """SDPA path == explicit softmax(QK^T) path with an empty controller (ptp_utils.py:238-253)."""
    from oracle import ref_loop
    g = torch.Generator().manual_seed(7)
    x, ctx = torch.randn((2, 4, 64, 64), generator=g), torch.randn((2, 77, 768), generator=g)
    with torch.no_grad():
        a = pipe.unet(x, torch.tensor(501), encoder_hidden_states=ctx)["sample"]
        with ref_loop.hooked(pipe.unet, ptp=None):
            b = pipe.unet(x, torch.tensor(501), encoder_hidden_states=ctx)["sample"]
    assert (a - b).abs().max().item() < 1e-4


def test_ddim_round_trip_and_direction_consistency():
    from oracle import ref_loop, sd15
    sch = ref_loop.Ddim(50)
    g = torch.Generator().manual_seed(3)
    x, e = torch.randn((1, 4, 64, 64), generator=g), torch.randn((1, 4, 64, 64), generator=g)
    for t in (981, 501, 21):
        fwd = sch.inverse(e, t, x)            # t-20 -> t
        back = sch.step(e, t, fwd)            # t -> t-20 with the same eps at eta = 0
        assert (back - x).abs().max().item() < 1e-5
    d = sd15.sd_scheduler()
    d.set_timesteps(50)
    assert torch.equal(d.timesteps, sch.timesteps)
    out = d.step(e, 981, x, eta=0.3, variance_noise=torch.ones_like(x)).prev_sample
    ref = sch.step(e, 981, x, eta=0.3, noise=torch.ones_like(x), add_noise=True)
    assert torch.equal(out, ref)


def test_ref_loop_reproduces_reference_trajectory(pipe):
    """etainv + ptp refine (3 steps): oracle/ref_loop.py vs the golden written by the reference's own code."""
    from eta_inversion_b200 import synthetic as syn
    from oracle import ref_loop
    from oracle.run_reference import SCENARIOS, SRC, TGT
    name = "etainv_ptp_refine_3"
    inv_kw, ed, _, cfg, inv_cfg = SCENARIOS[name]
    gold = np.load(GOLDEN / f"{name}.npz")
    out = ref_loop.edit(pipe, syn.synthetic_image(0), SRC, TGT, inverter="etainv", editor=ed, steps=3, ptp_cfg=cfg,
                        eta=inv_kw["eta"], decode=False)
    assert (out["inv_latents"] - torch.from_numpy(gold["inv_latents"])).abs().max().item() < 1e-4
    assert (out["bwd_latents"] - torch.from_numpy(gold["bwd_latents"])).abs().max().item() < 1e-4
    assert out["picks"] == gold["picks"].tolist()
    assert (out["fwd_mean"][1] - torch.from_numpy(gold["fwd_mean_map"])).abs().max().item() < 1e-5


@pytest.mark.parametrize("name", ["etainv_simple_3"])
def test_ref_loop_is_bit_exact_on_more_reference_goldens(pipe, name):
    """The self-contained port (what `bench.py --impl reference` and smoke() use as the CPU checker) against goldens the
    reference's own classes wrote for other inverter x editor pairs.  (diffinv+simple, dirinv+ptp and dirinv+masactrl were
    checked the same way when the goldens were generated: max-abs 0.0; one is kept here to bound the CPU suite's run time.)"""
    from eta_inversion_b200 import synthetic as syn
    from oracle import ref_loop
    from oracle.run_reference import SCENARIOS, SRC, TGT
    inv_kw, ed, _, cfg, _ = SCENARIOS[name]
    gold = np.load(GOLDEN / f"{name}.npz")
    kw = dict(inverter=inv_kw["type"], editor=ed, steps=inv_kw["num_inference_steps"], ptp_cfg=cfg, decode=False)
    if "eta" in inv_kw:
        kw["eta"] = inv_kw["eta"]
    with torch.no_grad():
        out = ref_loop.edit(pipe, syn.synthetic_image(0), SRC, TGT, **kw)
    assert torch.equal(out["inv_latents"], torch.from_numpy(gold["inv_latents"]))
    assert torch.equal(out["bwd_latents"], torch.from_numpy(gold["bwd_latents"]))
    if "picks" in gold.files:
        assert out["picks"] == gold["picks"].tolist()
