#!/usr/bin/env python
"""Edit a single image on the B200 engine.  Same flags as the reference CLI (edit_image.py:133-149); differences:
--guidance_scale_* accept floats, --edit_cfg really loads the YAML, --prec also takes bf16, and
--model defaults to the synthetic SD-1.5-architecture model because no checkpoint can be downloaded here."""
from __future__ import annotations

import argparse
import time
from pathlib import Path
from typing import List, Optional, Tuple

import torch

import eta_inversion_b200 as etai
from eta_inversion_b200.inversion.diffusion_inversion import DiffusionInversion


def split_to_words(prompt: str) -> List[str]:
    return (prompt[:-1] if prompt.endswith(".") else prompt).split(" ")


def get_edit_word(source_prompt: str, target_prompt: str) -> Optional[Tuple[str, str]]:
    s, t = split_to_words(source_prompt), split_to_words(target_prompt)
    if len(s) != len(t):
        return None
    diffs = [(a, b) for a, b in zip(s, t) if a != b]
    return diffs[0] if len(diffs) == 1 else None


def default_ptp_cfg(source_prompt: str, target_prompt: str):
    """Default prompt-to-prompt config from the single differing word (edit_image.py:77-102 of the reference)."""
    w = get_edit_word(source_prompt, target_prompt)
    if w is None:
        return None
    return dict(is_replace_controller=False, prompts=[source_prompt, target_prompt], cross_replace_steps={'default_': .4},
                self_replace_steps=0.6, blend_words=((w[0],), (w[1],)), equilizer_params={"words": (w[1],), "values": (2,)})


@torch.no_grad()
def main(input: str, model: str, source_prompt: str, target_prompt: str, output: Optional[str], inv_method: Optional[str],
         edit_method: Optional[str], scheduler: Optional[str], steps: Optional[int], guidance_scale_bwd: Optional[float],
         guidance_scale_fwd: Optional[float], edit_cfg: Optional[str], prec: Optional[str]) -> None:
    import cv2
    inp = Path(input)
    output = output or str(inp.parent / (inp.name + "_inv" + inp.suffix))
    pipe, (preproc, postproc) = etai.load_diffusion_model(model, "cuda", variant=prec)
    inv_method, edit_method = inv_method or "etainv", edit_method or "ptp"
    if edit_cfg is not None:
        import yaml
        cfg = yaml.safe_load(Path(edit_cfg).read_text())
    elif edit_method in ("ptp", "etaedit"):
        cfg = default_ptp_cfg(source_prompt, target_prompt)
        if cfg is None:
            print("Provide a edit_cfg for prompt-to-prompt if source and target prompt differ in more than one word.")
            return
        print(f"Using default ptp config:\n{cfg}")
    else:
        cfg = None
    inverter = etai.load_inverter(model=pipe, type=inv_method, scheduler=scheduler, num_inference_steps=steps,
                                  guidance_scale_bwd=guidance_scale_bwd, guidance_scale_fwd=guidance_scale_fwd)
    editor = etai.load_editor(inverter=inverter, type=edit_method)
    image = preproc(inp)
    idx = next((i for i, (s, t) in enumerate(zip(source_prompt.split(" "), target_prompt.split(" "))) if s != t), None)
    torch.cuda.synchronize()
    t1 = time.time()
    res = editor.edit(image, source_prompt, target_prompt, cfg=cfg, inv_cfg=dict(edit_word_idx=(idx, idx)))
    torch.cuda.synchronize()
    t2 = time.time()
    cv2.imwrite(output, cv2.cvtColor(postproc(res["image"]), cv2.COLOR_RGB2BGR))
    if "image_inv" in res:
        o = Path(output)
        cv2.imwrite(str(o.parent / (o.stem + "_inv" + o.suffix)), cv2.cvtColor(postproc(res["image_inv"]), cv2.COLOR_RGB2BGR))
    print(f"Saved result to {output}")
    print(f"Took {t2 - t1}s")


def parse_args():
    p = argparse.ArgumentParser(formatter_class=argparse.RawTextHelpFormatter, description="Edits a single image.")
    p.add_argument("--input", required=True, help="Path to image to invert.")
    p.add_argument("--model", default="synthetic-sd15", help="Diffusion Model.")
    p.add_argument("--source_prompt", required=True, help="Prompt to use for inversion.")
    p.add_argument("--target_prompt", required=True, help="Prompt to use for inversion.")
    p.add_argument("--output", help="Path for output image.")
    p.add_argument("--inv_method", choices=etai.get_inversion_methods(), help="Inversion method.")
    p.add_argument("--edit_method", choices=etai.get_edit_methods(), help="Editing method.")
    p.add_argument("--edit_cfg", help="Path to yaml file for editor configuration. Often needed for prompt-to-prompt.")
    p.add_argument("--scheduler", help="Which scheduler to use.", choices=DiffusionInversion.get_available_schedulers())
    p.add_argument("--steps", type=int, help="How many diffusion steps to use.")
    p.add_argument("--guidance_scale_bwd", type=float, help="Classifier free guidance scale for backward diffusion (denoising).")
    p.add_argument("--guidance_scale_fwd", type=float, help="Classifier free guidance scale for forward diffusion (inversion).")
    p.add_argument("--prec", choices=["fp16", "bf16", "fp32"], help="Precision for diffusion.")
    return vars(p.parse_args())


if __name__ == "__main__":
    main(**parse_args())
