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


NOT_BUILT = {}  # scenario name -> reason (kept so the parametrised list always equals the golden list)
NON_DEGENERATE_ETA = ("etainv_ptp_replace_5", "etainv_ptp_refine_3", "etainv_pnp_4", "etainv_simple_3",
                      "etainv_ptp_replace_bwdmask_3", "etainv_ptp_replace_gtmask_3")
FULL_SIZE = ("diffinv_simple_10", "etainv_ptp_replace_50", "etainv_masactrl_50")  # BASELINE.json configs 1-3 at their step count


def scenario_inputs(name):
    inv_kw, ed_type, ed_kw, cfg, inv_cfg = SCENARIOS[name]
    if inv_kw["type"] in ("etainv", "ddpminv", "cyclediff"):
        inv_kw = {**inv_kw, "noise_device": "cpu"}  # the goldens were written by the reference running on the CPU
    if inv_cfg is not None and isinstance(inv_cfg.get("mask"), str):
        from oracle.run_reference import gt_mask
        inv_cfg = {**inv_cfg, "mask": gt_mask()}
    return inv_kw, ed_type, ed_kw, cfg, inv_cfg


def run_scenario(pipe, name):
    import eta_inversion_b200 as etai
    from eta_inversion_b200 import synthetic as syn
    inv_kw, ed_type, ed_kw, cfg, inv_cfg = scenario_inputs(name)
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


def _psnr_gate(variant: str, name: str) -> float:
    """fp16: 35 dB (BASELINE.json north_star).  bf16: the synthetic model's peaky cross-attention (logit gain 4x) and 8 mantissa
    bits do not allow 35 dB -- the REFERENCE's own loops with the oracle model cast to bf16 on the CPU reach 27..34 dB against
    their fp32 image (tests/golden/ref_16bit_psnr.json, written by scripts/ref_16bit_psnr.py).  The bf16 gate is therefore
    'no worse than the reference's bf16 arithmetic' where that was measured, and a 28 dB floor elsewhere."""
    if variant != "bf16":
        return 35.0
    import json
    ref = json.loads((GOLDEN / "ref_16bit_psnr.json").read_text()).get("bf16", {}).get(name)
    return min(35.0, ref - 0.5) if ref is not None else 28.0


_FP32_IMAGES = {}  # scenario -> (image, image_inv) of the fp32 engine, the reference of the 16-bit PSNR gates


def _picks(inverter):
    return [int(p.item()) for p in getattr(inverter, "picks", [])]


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_fp32_trajectory_matches_reference(pipe_fp32, name):
    if name in NOT_BUILT:
        pytest.skip(NOT_BUILT[name])
    gold = np.load(GOLDEN / f"{name}.npz")
    res, rec, inverter = run_scenario(pipe_fp32, name)
    _FP32_IMAGES[name] = (res["image"].float().cpu(), res["image_inv"].float().cpu())
    if "nti_latents" in gold.files:  # null-text inversion also steps (one source row) inside its optimisation loop
        nti = torch.stack([l.cpu() for l in rec["bwd"] if l.shape != rec["bwd"][-1].shape])
        rec["bwd"] = [l for l in rec["bwd"] if l.shape == rec["bwd"][-1].shape]
        err_nti = (nti - torch.from_numpy(gold["nti_latents"])).abs().amax(dim=(1, 2, 3, 4))
        print(f"{name}: latents after each optimised null-text step max-abs {err_nti.tolist()}; inner steps "
              f"{inverter.inner_steps_taken}")
        assert err_nti.max() < TOL_LATENT
    inv = torch.stack([_t(l).cpu() for l in rec["inv"]["latents"]])[torch.from_numpy(gold["inv_steps_kept"])]
    err_inv = (inv - torch.from_numpy(gold["inv_latents"])).abs().amax(dim=(1, 2, 3, 4))
    bwd = torch.stack([l.cpu() for l in rec["bwd"]])[torch.from_numpy(gold["bwd_steps_kept"])]
    err_bwd = (bwd - torch.from_numpy(gold["bwd_latents"])).abs().amax(dim=(1, 2, 3, 4))
    print(f"{name}: inversion per-step max-abs {err_inv.tolist()}\n{name}: denoise per-step max-abs {err_bwd.tolist()}")
    if "fwd_mean_map" in gold.files:
        m = inverter.attn_maps_forward["mean"][SCENARIOS[name][4]["edit_word_idx"][0]].cpu()
        frac = float((m > 0.2).float().mean())
        print(f"{name}: fwd_mean map max-abs {(m - torch.from_numpy(gold['fwd_mean_map'])).abs().max():.2e}, "
              f"eta-mask coverage {frac:.3f} (golden {float(gold['eta_mask_fraction']):.3f})")
        assert (m - torch.from_numpy(gold["fwd_mean_map"])).abs().max() < 1e-3
        # the masked-eta path must not run in its degenerate all-ones form ("What's weak" #2).  The mean map
        # over many steps of the random-init model flattens out, so the long scenarios (>= 6 steps) cover ~everything;
        # NON_DEGENERATE_ETA lists the scenarios whose reference mask is mixed and test_eta_masks_are_mixed pins the list.
        if name in NON_DEGENERATE_ETA:
            assert 0.05 < float(gold["eta_mask_fraction"]) < 0.95 and 0.05 < frac < 0.95
        assert abs(frac - float(gold["eta_mask_fraction"])) < 2e-3
    if "picks" in gold.files:
        picks = _picks(inverter)
        print(f"{name}: noise picks {picks} (reference {gold['picks'].tolist()})")
        assert picks == gold["picks"].tolist()
    if "uncond_embeddings" in gold.files:
        # Adam divides by sqrt(v): an element whose gradient is at the rounding-noise level still moves by ~lr per step, in a
        # direction the noise decides, so single elements may differ by a few lr (1e-2) although they do not matter for the
        # output (the latents above are gated at 1e-3).  Gate the optimisation itself on the direction and size of the
        # update, and bound the outliers.
        u = torch.stack([x.float().cpu() for x in rec["inv"]["uncond_embeddings"]])
        ug = torch.from_numpy(gold["uncond_embeddings"])
        u0 = rec["inv"]["context"][:1].float().cpu()
        du, dug = (u - u0).flatten(1), (ug - u0).flatten(1)
        cos = torch.nn.functional.cosine_similarity(du, dug, dim=1)
        frac_off = ((u - ug).abs() > 1e-3).float().mean().item()
        print(f"{name}: optimised null-text embeddings: update cosine per step {cos.tolist()}, |update| ratio "
              f"{(du.norm(dim=1) / dug.norm(dim=1)).tolist()}, elements off by > 1e-3: {100 * frac_off:.2f} %, "
              f"max-abs {(u - ug).abs().max():.2e}")
        assert cos.min() > 0.99 and frac_off < 0.05 and (u - ug).abs().max() < 3.5e-2
    assert err_inv.max() < TOL_LATENT
    assert err_bwd.max() < TOL_LATENT
    assert (_t(res["latent"]).cpu() - torch.from_numpy(gold["latent"])).abs().max() < TOL_LATENT
    assert (_t(res["latent_inv"]).cpu() - torch.from_numpy(gold["latent_inv"])).abs().max() < TOL_LATENT
    pooled = torch.nn.functional.avg_pool2d(res["image"].float().cpu(), 8)
    assert (pooled - torch.from_numpy(gold["image_pool8"])).abs().max() < 5e-3
    if "image_f16" in gold.files:
        from eta_inversion_b200.metrics import psnr
        p = psnr(res["image"].float().cpu(), torch.from_numpy(gold["image_f16"]).float())
        print(f"{name}: fp32 engine image vs reference image PSNR {p:.1f} dB")
        assert p >= 50.0


def test_eta_masks_are_mixed():
    """At least six loop goldens run the masked eta step with a genuinely mixed mask (5 % .. 95 % of the latent pixels)."""
    mixed = [n for n in SCENARIOS if (GOLDEN / f"{n}.npz").exists() and "eta_mask_fraction" in np.load(GOLDEN / f"{n}.npz").files
             and 0.05 < float(np.load(GOLDEN / f"{n}.npz")["eta_mask_fraction"]) < 0.95]
    assert set(mixed) == set(NON_DEGENERATE_ETA) and len(mixed) >= 6, mixed


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


@pytest.fixture(scope="module", params=["fp16", "bf16"])
def pipe_16(request):
    import eta_inversion_b200 as etai
    pipe, _ = etai.load_diffusion_model("synthetic-sd15", "cuda", variant=request.param)
    pipe.variant = request.param
    yield pipe
    pipe.unet.close()


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_16bit_psnr_gate(pipe_fp32, pipe_16, name):
    """fp16 / bf16 mode (the tcgen05 kernels): final decoded image PSNR >= the gate of _psnr_gate (35 dB of BASELINE.json's
    north_star in fp16; in bf16 what the reference's own arithmetic reaches on this model, see there) for EVERY
    scenario -- replace, refine + reweight (the blend_a branch of the fused cross-attention), MasaCtrl K/V remap,
    plug-and-play injection, the other inverters, and configs 1-3 at their full step count -- against the fp32 engine
    image (itself pinned to the reference per step by the test above) and, where the golden holds the reference's full
    image, against the reference image directly.  Noise-pick agreement with the reference is reported."""
    if name in NOT_BUILT:
        pytest.skip(NOT_BUILT[name])
    from eta_inversion_b200.metrics import psnr
    if name not in _FP32_IMAGES:
        ref, _, _ = run_scenario(pipe_fp32, name)
        _FP32_IMAGES[name] = (ref["image"].float().cpu(), ref["image_inv"].float().cpu())
    img32, inv32 = _FP32_IMAGES[name]
    out, _, inverter = run_scenario(pipe_16, name)
    p_edit, p_inv = psnr(out["image"].float().cpu(), img32), psnr(out["image_inv"].float().cpu(), inv32)
    gold = np.load(GOLDEN / f"{name}.npz")
    msg = f"{name} [{pipe_16.variant}] vs fp32 engine: edit {p_edit:.1f} dB, reconstruction {p_inv:.1f} dB"
    if "image_f16" in gold.files:
        p_ref = psnr(out["image"].float().cpu(), torch.from_numpy(gold["image_f16"]).float())
        msg += f"; vs reference image {p_ref:.1f} dB"
        assert p_ref >= _psnr_gate(pipe_16.variant, name), msg
    if "picks" in gold.files:
        picks, want = _picks(inverter), gold["picks"].tolist()
        agree = sum(a == b for a, b in zip(picks, want))
        msg += f"; noise picks agree {agree}/{len(want)}"
    print(msg)
    gate = _psnr_gate(pipe_16.variant, name)
    assert p_edit >= gate and p_inv >= (gate if pipe_16.variant == "fp16" else min(gate, 30.0)), msg + f" (gate {gate:.1f} dB)"


def test_config2_as_benchmarked_matches_reference():
    """BASELINE config 2 exactly as bench.py runs it -- fp16, 50 + 50 steps, 4 edits in lock step (B = 8 / 16), two
    groups in flight on two engine handles that share one copy of the weights -- against the reference's golden image
    of the 50-step edit (lane 0 of every group edits the golden's image / prompts)."""
    import eta_inversion_b200 as etai
    from eta_inversion_b200 import synthetic as syn
    from eta_inversion_b200.batching import run_pipelined
    from eta_inversion_b200.metrics import psnr
    from eta_inversion_b200.models import clone_pipeline
    name = "etainv_ptp_replace_50"
    gold = np.load(GOLDEN / f"{name}.npz")
    inv_kw, ed_type, ed_kw, cfg, inv_cfg = scenario_inputs(name)
    pipe, _ = etai.load_diffusion_model("synthetic-sd15", "cuda", variant="fp16", max_batch=16)
    pipes = [pipe, clone_pipeline(pipe)]
    inverters = []

    def make_editor(p):
        inv = etai.load_inverter(model=p, **inv_kw)
        inverters.append(inv)
        return etai.load_editor(inverter=inv, type=ed_type, **ed_kw)
    others = [("a dog on a bench", "a fox on a bench", "dog", "fox"), ("a house by a lake", "a castle by a lake", "house", "castle"),
              ("a car in the rain", "a boat in the rain", "car", "boat")]

    def group(g):
        jobs = [dict(image=syn.synthetic_image(0).cuda(), source_prompt=SRC, target_prompt=TGT, cfg={**cfg}, inv_cfg=dict(inv_cfg))]
        for i, (s_, t_, a, b) in enumerate(others):
            jobs.append(dict(image=syn.synthetic_image(10 * g + i + 1).cuda(), source_prompt=s_, target_prompt=t_,
                             cfg={**cfg, "blend_words": [[a], [b]], "equilizer_params": {"words": [b], "values": [2]}},
                             inv_cfg=dict(inv_cfg)))
        return jobs
    res = run_pipelined(pipes, [group(0), group(1)], make_editor)
    ref_img = torch.from_numpy(gold["image_f16"]).float()
    for g in range(2):
        out = res[g][0]
        p = psnr(out["image"].float().cpu(), ref_img)
        lat_err = (out["latent"].float().cpu() - torch.from_numpy(gold["latent"])).abs().max().item()
        print(f"config 2 as benchmarked, group {g} lane 0: image PSNR vs reference {p:.1f} dB, final latent max-abs {lat_err:.3e}")
        assert p >= 35.0
    # both groups ran the same edit in lane 0 on different engine handles: same weights, same result
    assert psnr(res[0][0]["image"].float().cpu(), res[1][0]["image"].float().cpu()) >= 60.0
    want = gold["picks"].tolist()
    for inv in inverters[::4][:2]:  # lane 0 of each group
        picks = _picks(inv)
        print(f"noise picks agree with the reference at {sum(a == b for a, b in zip(picks, want))}/{len(want)} steps")
    for p_ in pipes:
        p_.unet.close()


def test_masactrl_8_cobatched_pairs_equal_sequential():
    """BASELINE config 3 (batch of 8 prompt pairs): 8 etainv + MasaCtrl edits in lock step (B = 16 inversion, B = 32 edit,
    mutual self-attention K/V remap shifted per lane) reproduce the one-at-a-time results on the fp32 path."""
    import eta_inversion_b200 as etai
    from eta_inversion_b200 import synthetic as syn
    from eta_inversion_b200.batching import run_lockstep
    pipe, _ = etai.load_diffusion_model("synthetic-sd15", "cuda", variant="fp32", max_batch=32)
    words = ["cat", "tiger", "dog", "fox", "house", "castle", "car", "boat", "tree"]

    def make_editor(p):
        inv = etai.load_inverter(type="etainv", model=p, scheduler="ddim", num_inference_steps=6, noise_device="cpu")
        return etai.load_editor(type="masactrl", inverter=inv)
    jobs = [dict(image=syn.synthetic_image(i).cuda(), source_prompt=f"a {words[i]} sitting next to a mirror",
                 target_prompt=f"a {words[i + 1]} sitting next to a mirror", inv_cfg=dict(edit_word_idx=(1, 1))) for i in range(8)]
    par = run_lockstep(pipe, [dict(j) for j in jobs], make_editor)
    with torch.no_grad():
        seq = [make_editor(pipe).edit(**dict(jobs[i])) for i in (0, 3, 7)]
    for i, a in zip((0, 3, 7), seq):
        err = (a["latent"] - par[i]["latent"]).abs().max().item()
        print(f"masactrl lane {i} of 8: lock-step vs sequential latent max-abs {err:.2e}")
        assert err < 1e-5
        assert (a["latent_inv"] - par[i]["latent_inv"]).abs().max().item() < 1e-5
    pipe.unet.close()


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


def test_lockstep_pnp_equals_sequential():
    """Plug-and-Play edits in lock step: the merged batch is laid out role-major so that the engine's feature injection
    (rows [0, n) copied over rows [n, 3n), pnp_utils.py:172-177) and the q/k remap serve every lane with one forward; three
    etainv + pnp edits (B = 6 inversion... B = 9 edit forwards) reproduce the one-at-a-time results on the fp32 path."""
    import eta_inversion_b200 as etai
    from eta_inversion_b200 import synthetic as syn
    from eta_inversion_b200.batching import run_lockstep
    pipe, _ = etai.load_diffusion_model("synthetic-sd15", "cuda", variant="fp32", max_batch=12)

    def make_editor(p):
        inv = etai.load_inverter(type="etainv", model=p, scheduler="ddim", num_inference_steps=4, noise_device="cpu")
        return etai.load_editor(type="pnp", inverter=inv)
    words = ["cat", "tiger", "dog", "fox"]
    jobs = [dict(image=syn.synthetic_image(i).cuda(), source_prompt=f"a {words[i]} sitting next to a mirror",
                 target_prompt=f"a {words[i + 1]} sitting next to a mirror", inv_cfg=dict(edit_word_idx=(1, 1))) for i in range(3)]
    with torch.no_grad():
        seq = [make_editor(pipe).edit(**dict(j)) for j in jobs]
    par = run_lockstep(pipe, [dict(j) for j in jobs], make_editor)
    for i, (a, b) in enumerate(zip(seq, par)):
        err = (a["latent"] - b["latent"]).abs().max().item()
        print(f"pnp lane {i} of 3: lock-step vs sequential latent max-abs {err:.2e}")
        assert err < 1e-5
        assert (a["image"].float() - b["image"].float()).abs().max().item() < 1e-4
    pipe.unet.close()
