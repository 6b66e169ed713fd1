"""Wall-clock phases of one etainv+ptp edit (synchronised at the phase boundaries): where the non-UNet time goes.
    python scripts/phase_times.py"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import eta_inversion_b200 as etai  # noqa: E402
from eta_inversion_b200 import synthetic as syn  # noqa: E402

cfg = dict(is_replace_controller=True, cross_replace_steps={"default_": .8}, self_replace_steps=.5,
           blend_words=[["cat"], ["tiger"]], equilizer_params={"words": ["tiger"], "values": [2]})
pipe, _ = etai.load_diffusion_model("synthetic-sd15", "cuda:0", variant="fp16")
inv = etai.load_inverter(type="etainv", model=pipe, scheduler="ddim", num_inference_steps=50)
ed = etai.load_editor(type="ptp", inverter=inv)
img = syn.synthetic_image(0).cuda()
SRC, TGT = "a cat sitting next to a mirror", "a tiger sitting next to a mirror"


def tick(t, name, acc):
    torch.cuda.synchronize()
    now = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + 1e3 * (now - t)
    return now


def one(acc):
    with torch.no_grad():
        t = time.perf_counter()
        src, tgt = inv.create_context(SRC), inv.create_context(TGT)
        t = tick(t, "contexts", acc)
        lat = inv.image2latent(img) if hasattr(inv, "image2latent") else None
        t = tick(t, "vae_encode(extra)", acc)
        res = inv.invert(img, prompt=SRC, context=src, inv_cfg=dict(edit_word_idx=(1, 1)))
        t = tick(t, "invert(total)", acc)
        ctrl = ed.make_controller(image=img, source_prompt=SRC, target_prompt=TGT, inv_res=res, **{**cfg})
        t = tick(t, "make_controller", acc)
        with inv.use_controller(ctrl):
            out = inv.sample(res, context=[src, tgt])
        t = tick(t, "sample(total)", acc)
        l = out["latent"] if isinstance(out, dict) and "latent" in out else None
        if l is not None:
            inv.latent2image(l) if hasattr(inv, "latent2image") else None
            t = tick(t, "vae_decode(extra)", acc)


for _ in range(2):
    one({})
acc = {}
N = 3
for _ in range(N):
    one(acc)
for k, v in acc.items():
    print(f"{k:24s} {v / N:8.1f} ms")
