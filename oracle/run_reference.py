"""ORACLE tooling (test infrastructure): run the UNMODIFIED reference loop/controller code
(/root/reference/modules/**) on CPU fp32 over the diffusers restatement in oracle/sd15.py and
write golden fixtures to tests/golden/*.npz.

Only runs in the build container (needs /root/reference).  Nothing under tests/ -m gpu, smoke()
or bench.py reads /root/reference; they read the committed fixtures.

    python oracle/run_reference.py --scenario all

The process chdir()s to a scratch directory first and neuters ``os.system`` because importing
modules/inversion/eta_inversion.py executes ``rm -rf result/pie_eta_new/*`` (eta_inversion.py:19-20).
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
GOLDEN = REPO / "tests" / "golden"

SRC = "a cat sitting next to a mirror"
TGT = "a tiger sitting next to a mirror"
PTP_REPLACE = dict(is_replace_controller=True, cross_replace_steps={"default_": .8}, self_replace_steps=.5,
                   blend_words=[["cat"], ["tiger"]], equilizer_params={"words": ["tiger"], "values": [2]})  # test/test_edit.py:32-39
PTP_REFINE = dict(is_replace_controller=False, prompts=[SRC, TGT], cross_replace_steps={"default_": .4},
                  self_replace_steps=0.6, blend_words=((("cat",), ("tiger",))),
                  equilizer_params={"words": ("tiger",), "values": (2,)})  # edit_image.py:87-98

SCENARIOS = {
    # name: (inverter kwargs, editor type, editor kwargs, edit cfg, inv_cfg)
    "diffinv_simple_4": (dict(type="diffinv", scheduler="ddim", num_inference_steps=4), "simple", {}, None, None),
    "etainv_ptp_replace_5": (dict(type="etainv", scheduler="ddim", num_inference_steps=5, eta=(0.0, 0.4)), "ptp", {},
                             PTP_REPLACE, dict(edit_word_idx=(1, 1))),
    "etainv_ptp_refine_3": (dict(type="etainv", scheduler="ddim", num_inference_steps=3,
                                 eta=[[0.6, 0.0], [1.0, 0.7]]), "ptp", {}, PTP_REFINE, dict(edit_word_idx=(1, 1))),
    "etainv_masactrl_6": (dict(type="etainv", scheduler="ddim", num_inference_steps=6, eta=(0.0, 0.4)), "masactrl", {},
                          None, dict(edit_word_idx=(1, 1))),
    "dirinv_ptp_replace_3": (dict(type="dirinv", scheduler="ddim", num_inference_steps=3), "ptp", {}, PTP_REPLACE, None),
    "npi_simple_3": (dict(type="npi", scheduler="ddim", num_inference_steps=3), "simple", {}, None, None),
    "diffinv_pnp_5": (dict(type="diffinv", scheduler="ddim", num_inference_steps=5), "pnp", {}, None, None),
    "proxnpi_simple_3": (dict(type="proxnpi", scheduler="ddim", num_inference_steps=3), "simple", {}, None, None),
    # cross combinations of the registries (test/test_edit.py:66-108 sweeps inverters x editors)
    "npi_ptp_replace_3": (dict(type="npi", scheduler="ddim", num_inference_steps=3), "ptp", {}, PTP_REPLACE, None),
    "dirinv_masactrl_4": (dict(type="dirinv", scheduler="ddim", num_inference_steps=4), "masactrl", {}, None, None),
    "etainv_pnp_4": (dict(type="etainv", scheduler="ddim", num_inference_steps=4, eta=(0.0, 0.4)), "pnp", {}, None,
                     dict(edit_word_idx=(1, 1))),
    "etainv_simple_3": (dict(type="etainv", scheduler="ddim", num_inference_steps=3, eta=0.3), "simple", {}, None,
                        dict(edit_word_idx=(1, 1))),
    # DDPM inversion / CycleDiffusion: sampled forward trajectory + recovered noise maps, 36 % of the steps skipped
    "ddpminv_simple_6": (dict(type="ddpminv", num_inference_steps=6), "simple", {}, None, None),
    "cyclediff_ptp_replace_5": (dict(type="cyclediff", num_inference_steps=5), "ptp", {}, PTP_REPLACE, None),
    # DDIM inversion with pix2pix-zero's noise regularisation (auto-correlation + KL gradient steps on the noise)
    "regdiffinv_simple_3": (dict(type="regdiffinv", scheduler="ddim", num_inference_steps=3), "simple", {}, None, None),
    # EDICT: coupled latent pair, two UNet calls per step, exact inversion
    "edict_simple_4": (dict(type="edict", scheduler="ddim", num_inference_steps=4), "simple", {}, None, None),
    "edict_ptp_replace_3": (dict(type="edict", scheduler="ddim", num_inference_steps=3), "ptp", {}, PTP_REPLACE, None),
    # the other mask modes of EtaInversion.get_mask (eta_inversion.py:159-205): maps read from the edit loop's own controller
    # (bwd_source_target), direct-inversion leak to the target row (target_dirinv, masked by mask_dirinv), a soft mask
    # (thres None + pow), and a user-supplied ("gt") mask that is bilinearly resized to 64x64 (eta_inversion.py:285-286)
    "etainv_ptp_replace_bwdmask_3": (dict(type="etainv", scheduler="ddim", num_inference_steps=3, eta=(0.0, 0.4),
                                          mask_mode_cfg=dict(mask_eta="bwd_source_target", mask_dirinv="fwd_mean",
                                                             target_dirinv=0.5, thres=0.3)),
                                     "ptp", {}, PTP_REPLACE, dict(edit_word_idx=(1, 1))),
    "etainv_ptp_replace_gtmask_3": (dict(type="etainv", scheduler="ddim", num_inference_steps=3, eta=(0.0, 0.4),
                                         mask_mode_cfg=dict(mask_eta="gt", thres=None, pow=2.0)),
                                    "ptp", {}, PTP_REPLACE, dict(edit_word_idx=(1, 1), mask="GT_MASK")),
    # BASELINE.json configs at their stated size (per-step tensors stored every KEEP_EVERY-th step to keep the fixture small)
    "diffinv_simple_10": (dict(type="diffinv", scheduler="ddim", num_inference_steps=10), "simple", {}, None, None),
    "etainv_ptp_replace_50": (dict(type="etainv", scheduler="ddim", num_inference_steps=50, eta=(0.0, 0.4)), "ptp", {},
                              PTP_REPLACE, dict(edit_word_idx=(1, 1))),
    "etainv_masactrl_50": (dict(type="etainv", scheduler="ddim", num_inference_steps=50, eta=(0.0, 0.4)), "masactrl", {},
                           None, dict(edit_word_idx=(1, 1))),
    # null-text inversion (null_text_inversion.py:42-101): per-step Adam on the uncond embedding, autograd through the UNet
    "nti_ptp_replace_3": (dict(type="nti", scheduler="ddim", num_inference_steps=3, num_inner_steps=3), "ptp", {},
                          PTP_REPLACE, None),
}
FULL_IMAGE = ("etainv_ptp_replace_5", "etainv_ptp_replace_50", "etainv_masactrl_50", "diffinv_simple_10",
              "etainv_ptp_refine_3")
KEEP_EVERY = 5  # scenarios with more than 10 steps keep step 0, 5, 10, ... and the last one


def gt_mask() -> torch.Tensor:
    """User-supplied 512x512 soft mask of the 'gt' mask mode: a smooth blob, values in [0, 1]."""
    yy, xx = torch.meshgrid(torch.linspace(0, 1, 512), torch.linspace(0, 1, 512), indexing="ij")
    return torch.exp(-(((xx - 0.45) / 0.25) ** 2 + ((yy - 0.55) / 0.3) ** 2)).clamp(0, 1)


def keep_steps(n: int):
    if n <= 12:
        return list(range(n))
    return sorted(set(range(0, n, KEEP_EVERY)) | {n - 1})


def _setup_paths():
    os.system = lambda *a, **k: 0  # see module docstring
    os.chdir(tempfile.mkdtemp(prefix="etai_oracle_"))
    sys.path[:0] = [str(REPO / "oracle" / "shim"), str(REPO), str(REF)]


def build_model():
    from eta_inversion_b200 import synthetic as syn
    from oracle import sd15
    pipe = sd15.build_pipeline(syn.random_state_dict(syn.unet_param_spec(), 0),
                               syn.random_state_dict(syn.vae_param_spec(), 1), seed=0)
    return pipe, syn


def as_tensor(x) -> torch.Tensor:
    """EDICT carries a PAIR of coupled latents (a list): stored as one tensor with the pair stacked on the batch axis."""
    return torch.cat([as_tensor(v) for v in x]) if isinstance(x, (list, tuple)) else x


def pool8(img: torch.Tensor) -> np.ndarray:
    return torch.nn.functional.avg_pool2d(img, 8).numpy()


def run_unet_fwd(pipe):
    g = torch.Generator().manual_seed(1234)
    x = torch.randn((4, 4, 64, 64), generator=g)
    ctx = torch.randn((4, 77, 768), generator=g)
    out = {}
    with torch.no_grad():
        for t in (981, 1):
            out[f"eps_t{t}"] = pipe.unet(x, torch.tensor(t), encoder_hidden_states=ctx)["sample"].numpy()
    np.savez_compressed(GOLDEN / "unet_fwd.npz", **out)
    print("unet_fwd", {k: float(np.abs(v).mean()) for k, v in out.items()})


def run_scenario(name, pipe, syn):
    import modules  # the reference's own package
    inv_kw, ed_type, ed_kw, cfg, inv_cfg = SCENARIOS[name]
    if inv_cfg is not None and isinstance(inv_cfg.get("mask"), str):
        inv_cfg = {**inv_cfg, "mask": gt_mask()}
    inverter = modules.load_inverter(model=pipe, **inv_kw)
    editor = modules.load_editor(inverter=inverter, type=ed_type, **ed_kw)
    image = syn.synthetic_image(0)
    rec = {"bwd_latents": [], "bwd_eps": [], "picks": []}

    # record every backward step without touching the reference's code
    orig_psb = inverter.predict_step_backward

    def psb(*a, **k):
        new_latent, eps = orig_psb(*a, **k)
        rec["bwd_latents"].append(as_tensor(new_latent).detach().clone().numpy())
        if eps is not None:
            rec["bwd_eps"].append(eps.detach().clone().numpy())
        return new_latent, eps
    inverter.predict_step_backward = psb
    if hasattr(inverter, "get_eta_variance_noise"):
        orig_sample, orig_get = inverter.sample_variance_noise, inverter.get_eta_variance_noise
        last = {}

        def sample_vn(n, generator=None):
            last["c"] = orig_sample(n, generator)
            return last["c"]

        def get_eta(*a, **k):
            r = orig_get(*a, **k)
            idx = [i for i in range(last["c"].shape[0]) if torch.equal(last["c"][i], r["variance_noise"])]
            rec["picks"].append(idx[0])
            return r
        inverter.sample_variance_noise, inverter.get_eta_variance_noise = sample_vn, get_eta
    # LocalBlend mask coverage per call (ptp.py:18-29), to prove the goldens exercise a non-trivial mask
    from modules.utils import ptp as ref_ptp
    lb_frac = []
    orig_lb = ref_ptp.LocalBlend.get_mask

    def lb_get_mask(self, x_t, maps, alpha, use_pool):
        m = orig_lb(self, x_t, maps, alpha, use_pool)
        lb_frac.append(m.float().mean(dim=(1, 2, 3)).tolist())
        return m
    ref_ptp.LocalBlend.get_mask = lb_get_mask
    orig_invert = inverter.invert
    inv_box = {}

    def invert(*a, **k):
        inv_box["res"] = orig_invert(*a, **k)
        return inv_box["res"]
    inverter.invert = invert

    t0 = time.perf_counter()
    with torch.no_grad():  # (NullTextInversion re-enables grad around its inner optimisation itself)
        res = editor.edit(image, SRC, TGT, cfg=None if cfg is None else {**cfg}, inv_cfg=inv_cfg)
    dt = time.perf_counter() - t0
    ref_ptp.LocalBlend.get_mask = orig_lb
    inv_lat = [as_tensor(l).detach().numpy() for l in inv_box["res"]["latents"]]
    inv_eps = [e.detach().numpy() for e in inv_box["res"]["noise_preds"] if e is not None]
    # null-text inversion also calls predict_step_backward inside its optimisation (one source row); those steps are kept
    # apart from the edit loop's (source + target rows)
    nti_lat = [x for x in rec["bwd_latents"] if x.shape != rec["bwd_latents"][-1].shape]
    nti_eps = [x for x in rec["bwd_eps"] if x.shape != rec["bwd_eps"][-1].shape]
    rec["bwd_latents"] = [x for x in rec["bwd_latents"] if x.shape == rec["bwd_latents"][-1].shape]
    rec["bwd_eps"] = [x for x in rec["bwd_eps"] if x.shape == rec["bwd_eps"][-1].shape]
    ki, kb = keep_steps(len(inv_lat)), keep_steps(len(rec["bwd_latents"]))
    out = dict(
        inv_latents=np.stack([inv_lat[i] for i in ki]), inv_steps_kept=np.array(ki),
        inv_eps=np.stack([inv_eps[i] for i in keep_steps(len(inv_eps))] or [np.zeros(0)]),
        bwd_latents=np.stack([rec["bwd_latents"][i] for i in kb]), bwd_steps_kept=np.array(kb),
        bwd_eps=np.stack([rec["bwd_eps"][i] for i in keep_steps(len(rec["bwd_eps"]))] or [np.zeros(0)]),
        latent=as_tensor(res["latent"]).numpy(), latent_inv=as_tensor(res["latent_inv"]).numpy(),
        image_pool8=pool8(res["image"]), image_inv_pool8=pool8(res["image_inv"]),
        image_mean=np.array([res["image"].mean().item(), res["image_inv"].mean().item()]),
        seconds=np.array(dt),
    )
    if nti_lat:
        out["nti_latents"], out["nti_eps"] = np.stack(nti_lat), np.stack(nti_eps)
    if lb_frac:
        out["localblend_mask_fraction"] = np.array(lb_frac)
    if name in FULL_IMAGE:  # PSNR gate of the 16-bit engine against the reference's fp32 image
        out["image_f16"] = res["image"].to(torch.float16).numpy()
    if rec["picks"]:
        out["picks"] = np.array(rec["picks"], dtype=np.int64)
    if getattr(inverter, "attn_maps_forward", None):
        out["fwd_mean_map"] = inverter.attn_maps_forward["mean"][inv_cfg["edit_word_idx"][0]].numpy()
        out["eta_mask_fraction"] = np.array(float((out["fwd_mean_map"] > 0.2).mean()))
    if "uncond_embeddings" in inv_box["res"]:  # null-text inversion: the optimised embedding of every step
        out["uncond_embeddings"] = np.stack([u.detach().numpy() for u in inv_box["res"]["uncond_embeddings"]])
    np.savez_compressed(GOLDEN / f"{name}.npz", **out)
    print(name, f"{dt:.1f}s", {k: v.shape for k, v in out.items()}, "picks", rec["picks"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", default="all")
    args = ap.parse_args()
    _setup_paths()
    torch.set_num_threads(os.cpu_count())
    GOLDEN.mkdir(parents=True, exist_ok=True)
    pipe, syn = build_model()
    names = ["unet_fwd"] + list(SCENARIOS) if args.scenario == "all" else args.scenario.split(",")
    for n in names:
        if n == "unet_fwd":
            run_unet_fwd(pipe)
        else:
            run_scenario(n, pipe, syn)


if __name__ == "__main__":
    main()
