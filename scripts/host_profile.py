"""Host-side (Python) profile of one etainv+ptp edit: where the CPU time between UNet forwards goes.
    python scripts/host_profile.py [--steps 50]"""
import argparse
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import eta_inversion_b200 as etai  # noqa: E402
from eta_inversion_b200 import synthetic as syn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--top", type=int, default=45)
args = ap.parse_args()
cfg = dict(is_replace_controller=True, cross_replace_steps={"default_": .8}, self_replace_steps=.5,
           blend_words=[["cat"], ["tiger"]], equilizer_params={"words": ["tiger"], "values": [2]})
pipe, _ = etai.load_diffusion_model("synthetic-sd15", "cuda:0", variant="fp16")
inv = etai.load_inverter(type="etainv", model=pipe, scheduler="ddim", num_inference_steps=args.steps)
ed = etai.load_editor(type="ptp", inverter=inv)
img = syn.synthetic_image(0).cuda()


def edit():
    with torch.no_grad():
        r = ed.edit(img, "a cat sitting next to a mirror", "a tiger sitting next to a mirror", cfg={**cfg},
                    inv_cfg=dict(edit_word_idx=(1, 1)))
    torch.cuda.synchronize()
    return r


edit(); edit()
t0 = time.perf_counter(); edit(); t1 = time.perf_counter()
print(f"edit wall {1e3 * (t1 - t0):.1f} ms")
pr = cProfile.Profile()
pr.enable(); edit(); pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(args.top)
