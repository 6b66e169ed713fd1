"""Host-side profile of ONE lock-step group (cobatch edits in lane threads, like bench.py): per-thread cProfile merged,
plus wall vs UNet-forward time.    python scripts/host_profile_group.py [--steps 50] [--cobatch 4]"""
import argparse
import cProfile
import pstats
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import eta_inversion_b200 as etai  # noqa: E402
from eta_inversion_b200 import synthetic as syn  # noqa: E402
from eta_inversion_b200.batching import run_lockstep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--cobatch", type=int, default=4)
ap.add_argument("--top", type=int, default=50)
args = ap.parse_args()
CB = args.cobatch
cfg = dict(is_replace_controller=True, cross_replace_steps={"default_": .8}, self_replace_steps=.5,
           blend_words=[["cat"], ["tiger"]], equilizer_params={"words": ["tiger"], "values": [2]})
pipe, _ = etai.load_diffusion_model("synthetic-sd15", "cuda:0", variant="fp16", max_batch=4 * CB)
pipe.cache_text_embeddings = False
imgs = [syn.synthetic_image(i).cuda() for i in range(CB)]
profs = []
CPU = []
PROFILE = [False]


class Wrapped:
    def __init__(self, ed):
        self.ed = ed

    def edit(self, **kw):
        import threading as _th
        c0 = time.thread_time()
        try:
            return self._edit(**kw)
        finally:
            CPU.append(time.thread_time() - c0)

    def _edit(self, **kw):
        import threading as _th
        if not PROFILE[0] or _th.current_thread().name != "etai-lane-0":  # one profiler may be active per process
            return self.ed.edit(**kw)
        pr = cProfile.Profile()
        profs.append(pr)
        pr.enable()
        try:
            return self.ed.edit(**kw)
        finally:
            pr.disable()


def make_editor(p):
    inv = etai.load_inverter(type="etainv", model=p, scheduler="ddim", num_inference_steps=args.steps, guidance_scale_bwd=7.5)
    return Wrapped(etai.load_editor(type="ptp", inverter=inv))


def group():
    jobs = [dict(image=im, source_prompt="a cat sitting next to a mirror", target_prompt="a tiger sitting next to a mirror",
                 cfg={**cfg}, inv_cfg=dict(edit_word_idx=(1, 1))) for im in imgs]
    r = run_lockstep(pipe, jobs, make_editor)
    torch.cuda.synchronize()
    return r


group(); group()
pipe.unet.time_forwards(True)
t0 = time.perf_counter(); group(); t1 = time.perf_counter()
unet_ms = pipe.unet.time_forwards(False)
print(f"group wall {1e3 * (t1 - t0):.1f} ms, UNet forwards {unet_ms:.1f} ms ({unet_ms / (1e3 * (t1 - t0)):.3f})")
print(f"CPU time (time.thread_time) of the lane threads for that group: {[round(1e3 * c) for c in CPU[-CB:]]} ms, sum "
      f"{1e3 * sum(CPU[-CB:]):.0f} ms -- all of it serialised by the GIL")
PROFILE[0] = True
t0 = time.perf_counter(); group(); t1 = time.perf_counter()
print(f"profiled group wall {1e3 * (t1 - t0):.1f} ms")
st = pstats.Stats(profs[0])
for p in profs[1:]:
    st.add(p)
st.sort_stats("tottime").print_stats(args.top)
