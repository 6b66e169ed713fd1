"""Build-container experiment (needs /root/reference): how peaky must the synthetic model's cross-attention be for the
eta mask (fwd_mean map > 0.2) and the LocalBlend mask (> 0.3) to be non-degenerate?

    python scripts/exp_peaky_attention.py --gain 1 2 3 4 --steps 3
"""
import argparse
import os
import sys
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from oracle import run_reference as rr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gain", type=float, nargs="+", default=[1.0])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--tok-gain", type=float, default=1.0, dest="tok_gain")
    ap.add_argument("--editor", default="ptp")
    ap.add_argument("--long", action="store_true")
    ap.add_argument("--vae-gain", type=float, default=1.0, dest="vae_gain")
    ap.add_argument("--image-noise", type=float, default=0.05, dest="image_noise")
    args = ap.parse_args()
    rr._setup_paths()
    torch.set_num_threads(os.cpu_count())
    from eta_inversion_b200 import synthetic as syn
    from oracle import sd15
    import modules
    from modules.utils import ptp as ref_ptp
    for gain in args.gain:
        syn.ATTN2_QK_GAIN = gain
        syn.TOKEN_EMB_GAIN = args.tok_gain
        syn.VAE_LATENT_GAIN = args.vae_gain
        syn.IMAGE_NOISE = args.image_noise
        pipe = sd15.build_pipeline(syn.random_state_dict(syn.unet_param_spec(), 0),
                                   syn.random_state_dict(syn.vae_param_spec(), 1), seed=0)
        inverter = modules.load_inverter(model=pipe, type="etainv", scheduler="ddim", num_inference_steps=args.steps,
                                         eta=(0.0, 0.4))
        editor = modules.load_editor(inverter=inverter, type=args.editor)
        masks = []
        orig = ref_ptp.LocalBlend.get_mask

        def get_mask(self, x_t, maps, alpha, use_pool):
            m = orig(self, x_t, maps, alpha, use_pool)
            masks.append(m.float().mean(dim=(1, 2, 3)).tolist())
            return m
        ref_ptp.LocalBlend.get_mask = get_mask
        t0 = time.perf_counter()
        cfg = {**rr.PTP_REPLACE} if args.editor == "ptp" else None
        src, tgt, wi = rr.SRC, rr.TGT, 1
        if args.long:
            src = "a photo of a small cat sitting quietly next to a large ornate golden mirror in a bright sunny room with a wooden floor"
            tgt = src.replace("cat", "tiger")
            wi = 5
        with torch.no_grad():
            res = editor.edit(syn.synthetic_image(0), src, tgt, cfg=cfg, inv_cfg=dict(edit_word_idx=(wi, wi)))
        ref_ptp.LocalBlend.get_mask = orig
        m = inverter.attn_maps_forward["mean"][wi]
        lat = res["latent"]
        q = torch.quantile(m.flatten(), torch.tensor([0.1, 0.25, 0.5, 0.75, 0.9]))
        print("   fwd_mean quantiles 10/25/50/75/90:", [round(v, 3) for v in q.tolist()])
        print(f"gain {gain}: fwd_mean min {m.min():.3f} mean {m.mean():.3f} frac>0.2 {(m > 0.2).float().mean():.3f} | "
              f"LocalBlend mask fractions (src,tgt rows) per step {masks} | latent absmax {lat.abs().max():.2f} "
              f"finite {bool(torch.isfinite(lat).all())} | {time.perf_counter() - t0:.0f}s", flush=True)


if __name__ == "__main__":
    main()
