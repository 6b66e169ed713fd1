"""Loop-level parity on the GPU: the engine's inverters/editors (through the C ABI) against golden trajectories that
the UNMODIFIED reference loop code produced on CPU fp32 over the oracle (oracle/run_reference.py), same seeded
weights / image / prompts / noise.  Gate (BASELINE.json north_star): every per-step latent within 1e-3 max-abs in
fp32; fp16 final image PSNR >= 35 dB."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.run_reference import SCENARIOS, SRC, TGT

pytestmark = pytest.mark.gpu

TOL_LATENT = 1e-3


@pytest.fixture(scope="module")
def pipe_fp32():
    import eta_inversion_b200 as etai
    pipe, _ = etai.load_diffusion_model("synthetic-sd15", "cuda", variant="fp32")
    return pipe


def _t(x):
    """EDICT carries a pair of coupled latents (a list): compared as one tensor, the pair stacked on the batch axis."""
    return torch.cat([_t(v) for v in x]) if isinstance(x, (list, tuple)) else x


def run_scenario(pipe, name):
    import eta_inversion_b200 as etai
    from eta_inversion_b200 import synthetic as syn
    inv_kw, ed_type, ed_kw, cfg, inv_cfg = SCENARIOS[name]
    if inv_kw["type"] in ("etainv", "ddpminv", "cyclediff"):
        inv_kw = {**inv_kw, "noise_device": "cpu"}  # the goldens were written by the reference running on the CPU
    inverter = etai.load_inverter(model=pipe, **inv_kw)
    editor = etai.load_editor(inverter=inverter, type=ed_type, **ed_kw)
    rec = {"bwd": [], "inv": None}
    orig_psb, orig_inv = inverter.predict_step_backward, inverter.invert

    def psb(*a, **k):
        out = orig_psb(*a, **k)
        rec["bwd"].append(_t(out[0]).detach().clone())
        return out

    def inv(*a, **k):
        rec["inv"] = orig_inv(*a, **k)
        return rec["inv"]
    inverter.predict_step_backward, inverter.invert = psb, inv
    with torch.no_grad():
        res = editor.edit(syn.synthetic_image(0).cuda(), SRC, TGT, cfg=None if cfg is None else {**cfg}, inv_cfg=inv_cfg)
    return res, rec, inverter


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_fp32_trajectory_matches_reference(pipe_fp32, name):
    gold = np.load(GOLDEN / f"{name}.npz")
    res, rec, inverter = run_scenario(pipe_fp32, name)
    inv = torch.stack([_t(l).cpu() for l in rec["inv"]["latents"]])
    err_inv = (inv - torch.from_numpy(gold["inv_latents"])).abs().amax(dim=(1, 2, 3, 4))
    bwd = torch.stack([l.cpu() for l in rec["bwd"]])
    err_bwd = (bwd - torch.from_numpy(gold["bwd_latents"])).abs().amax(dim=(1, 2, 3, 4))
    print(f"{name}: inversion per-step max-abs {err_inv.tolist()}\n{name}: denoise per-step max-abs {err_bwd.tolist()}")
    if "fwd_mean_map" in gold.files:
        m = inverter.attn_maps_forward["mean"][SCENARIOS[name][4]["edit_word_idx"][0]].cpu()
        print(f"{name}: fwd_mean map max-abs {(m - torch.from_numpy(gold['fwd_mean_map'])).abs().max():.2e}")
        assert (m - torch.from_numpy(gold["fwd_mean_map"])).abs().max() < 1e-3
    assert err_inv.max() < TOL_LATENT
    assert err_bwd.max() < TOL_LATENT
    assert (_t(res["latent"]).cpu() - torch.from_numpy(gold["latent"])).abs().max() < TOL_LATENT
    assert (_t(res["latent_inv"]).cpu() - torch.from_numpy(gold["latent_inv"])).abs().max() < TOL_LATENT
    pooled = torch.nn.functional.avg_pool2d(res["image"].float().cpu(), 8)
    assert (pooled - torch.from_numpy(gold["image_pool8"])).abs().max() < 5e-3


def test_state_does_not_leak_between_edits(pipe_fp32):
    """A / B / A (reference: test/test_edit.py:259-289): the second A run must reproduce the first bit-exactly."""
    import eta_inversion_b200 as etai
    from eta_inversion_b200 import synthetic as syn
    inverter = etai.load_inverter(model=pipe_fp32, type="etainv", scheduler="ddim", num_inference_steps=3)
    editor = etai.load_editor(inverter=inverter, type="ptp")
    img = syn.synthetic_image(0).cuda()
    cfg = SCENARIOS["etainv_ptp_replace_5"][3]
    with torch.no_grad():
        a1 = editor.edit(img, SRC, TGT, cfg={**cfg}, inv_cfg=dict(edit_word_idx=(1, 1)))
        editor.edit(syn.synthetic_image(3).cuda(), "a dog on a bench", "a fox on a bench",
                    cfg={**cfg, "blend_words": [["dog"], ["fox"]], "equilizer_params": {"words": ["fox"], "values": [2]}},
                    inv_cfg=dict(edit_word_idx=(1, 1)))
        a2 = editor.edit(img, SRC, TGT, cfg={**cfg}, inv_cfg=dict(edit_word_idx=(1, 1)))
    assert torch.equal(a1["latent"], a2["latent"]) and torch.equal(a1["image"], a2["image"])


def test_fp16_psnr_gate(pipe_fp32):
    """fp16/bf16 mode: final decoded image PSNR >= 35 dB against the fp32 result (which itself is pinned to the
    reference trajectory by the test above)."""
    import eta_inversion_b200 as etai
    from eta_inversion_b200.metrics import psnr
    name = "etainv_ptp_replace_5"
    ref, _, _ = run_scenario(pipe_fp32, name)
    pipe16, _ = etai.load_diffusion_model("synthetic-sd15", "cuda", variant="fp16")
    out, _, _ = run_scenario(pipe16, name)
    p_edit, p_inv = psnr(out["image"], ref["image"]), psnr(out["image_inv"], ref["image_inv"])
    gold = np.load(GOLDEN / f"{name}.npz")
    pooled = torch.nn.functional.avg_pool2d(out["image"].float().cpu(), 8)
    print(f"fp16 vs fp32 PSNR: edit {p_edit:.2f} dB, reconstruction {p_inv:.2f} dB; pooled max-abs vs reference "
          f"{(pooled - torch.from_numpy(gold['image_pool8'])).abs().max():.3e}")
    assert p_edit >= 35.0 and p_inv >= 35.0


def test_lockstep_cobatch_equals_sequential(pipe_fp32):
    """Two independent edits sharing every UNet forward (B=4 inversion, B=8 edit) reproduce the sequential results:
    the fp32 kernels are batch-invariant, so the comparison is tight."""
    import eta_inversion_b200 as etai
    from eta_inversion_b200 import synthetic as syn
    from eta_inversion_b200.batching import run_lockstep
    pipe, _ = etai.load_diffusion_model("synthetic-sd15", "cuda", variant="fp32", max_batch=8)
    cfg = SCENARIOS["etainv_ptp_replace_5"][3]

    def make_editor(p):
        inv = etai.load_inverter(type="etainv", model=p, scheduler="ddim", num_inference_steps=4)
        return etai.load_editor(type="ptp", inverter=inv)
    jobs = [dict(image=syn.synthetic_image(0).cuda(), source_prompt=SRC, target_prompt=TGT, cfg={**cfg},
                 inv_cfg=dict(edit_word_idx=(1, 1))),
            dict(image=syn.synthetic_image(5).cuda(), source_prompt="a dog on a bench", target_prompt="a fox on a bench",
                 cfg={**cfg, "blend_words": [["dog"], ["fox"]], "equilizer_params": {"words": ["fox"], "values": [2]}},
                 inv_cfg=dict(edit_word_idx=(1, 1)))]
    with torch.no_grad():
        seq = [make_editor(pipe).edit(**{**j, "cfg": {**j["cfg"]}}) for j in jobs]
    par = run_lockstep(pipe, [{**j, "cfg": {**j["cfg"]}} for j in jobs], make_editor)
    for a, b in zip(seq, par):
        err = (a["latent"] - b["latent"]).abs().max().item()
        print(f"lockstep vs sequential latent max-abs {err:.2e}")
        assert err < 1e-5
        assert (a["latent_inv"] - b["latent_inv"]).abs().max().item() < 1e-5
