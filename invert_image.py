#!/usr/bin/env python
"""Invert a single image on the B200 engine and save the reconstruction.  Same flags as the reference CLI
(invert_image.py:45-56: --input --prompt --output --method --scheduler --steps --guidance_scale_bwd/_fwd); additions:
--model (defaults to the synthetic SD-1.5-architecture model, no checkpoint can be downloaded here) and --prec;
--guidance_scale_* accept floats (the reference parses them as int)."""
from __future__ import annotations

import argparse
from pathlib import Path
from typing import Optional

import torch

import eta_inversion_b200 as etai
from eta_inversion_b200.inversion.diffusion_inversion import DiffusionInversion


@torch.no_grad()
def main(input: str, prompt: str, output: Optional[str], method: Optional[str], scheduler: Optional[str],
         steps: Optional[int], guidance_scale_bwd: Optional[float], guidance_scale_fwd: Optional[float],
         model: str = "synthetic-sd15", prec: Optional[str] = None) -> None:
    import cv2
    inp = Path(input)
    output = output or str(inp.parent / (inp.name + "_inv" + inp.suffix))  # invert_image.py:22-24
    pipe, (_, postproc) = etai.load_diffusion_model(model, "cuda", variant=prec)
    preproc = etai.StablePreprocess("cuda", size=512, center_crop=True, return_np=False, pil_resize=True)  # :30
    inverter = etai.load_inverter(model=pipe, type=method or "diffinv", scheduler=scheduler, num_inference_steps=steps,
                                  guidance_scale_bwd=guidance_scale_bwd, guidance_scale_fwd=guidance_scale_fwd)
    inv_res = inverter.invert_sample(preproc(inp), prompt)  # invert, then sample back with the same prompt (:37-38)
    cv2.imwrite(output, cv2.cvtColor(postproc(inv_res["image"]), cv2.COLOR_RGB2BGR))
    print(f"Saved result to {output}")


def parse_args():
    p = argparse.ArgumentParser(formatter_class=argparse.RawTextHelpFormatter, description="Inverts a single image.")
    p.add_argument("--input", required=True, help="Path to image to invert.")
    p.add_argument("--prompt", required=True, help="Prompt to use for inversion.")
    p.add_argument("--output", help="Path for output image.")
    p.add_argument("--method", choices=etai.get_inversion_methods(), help="Inversion method.")
    p.add_argument("--scheduler", help="Which scheduler to use.", choices=DiffusionInversion.get_available_schedulers())
    p.add_argument("--steps", type=int, help="How many diffusion steps to use.")
    p.add_argument("--guidance_scale_bwd", type=float, help="Classifier free guidance scale to use for backward diffusion (denoising).")
    p.add_argument("--guidance_scale_fwd", type=float, help="Classifier free guidance scale to use for forward diffusion (inversion).")
    p.add_argument("--model", default="synthetic-sd15", help="Diffusion Model.")
    p.add_argument("--prec", choices=["fp16", "bf16", "fp32"], help="Precision for diffusion.")
    return vars(p.parse_args())


if __name__ == "__main__":
    main(**parse_args())
