"""GPU busy/idle accounting for one co-batched step (torch.profiler, CUDA activities only).
    python scripts/timeline_cb.py --cobatch 4"""
import argparse
import sys
import time
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import eta_inversion_b200 as etai  # noqa: E402
from eta_inversion_b200 import synthetic as syn  # noqa: E402
from eta_inversion_b200.batching import run_lockstep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cobatch", type=int, default=4)
ap.add_argument("--steps", type=int, default=50)
args = ap.parse_args()
cfg = dict(is_replace_controller=True, cross_replace_steps={"default_": .8}, self_replace_steps=.5,
           blend_words=[["cat"], ["tiger"]], equilizer_params={"words": ["tiger"], "values": [2]})
pipe, _ = etai.load_diffusion_model("synthetic-sd15", "cuda:0", variant="fp16", max_batch=4 * args.cobatch)


def make_editor(lane_pipe):
    inv = etai.load_inverter(type="etainv", model=lane_pipe, scheduler="ddim", num_inference_steps=args.steps)
    return etai.load_editor(type="ptp", inverter=inv)


imgs = [syn.synthetic_image(i).cuda() for i in range(args.cobatch)]


def step():
    jobs = [dict(image=im, source_prompt="a cat sitting next to a mirror", target_prompt="a tiger sitting next to a mirror",
                 cfg={**cfg}, inv_cfg=dict(edit_word_idx=(1, 1))) for im in imgs]
    r = run_lockstep(pipe, jobs, make_editor)
    torch.cuda.synchronize()
    return r


for _ in range(3):
    step()
t0 = time.perf_counter(); step(); wall = 1e3 * (time.perf_counter() - t0)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter(); step(); wall_p = 1e3 * (time.perf_counter() - t0)
ka = prof.key_averages()
tot = sum(e.self_device_time_total for e in ka) / 1e3
print(f"wall {wall:.1f} ms (profiled {wall_p:.1f} ms), sum of GPU kernel time {tot:.1f} ms")
etai_t = sum(e.self_device_time_total for e in ka if "etai" in e.key) / 1e3
print(f"engine kernels {etai_t:.1f} ms, other {tot - etai_t:.1f} ms")
rows = sorted(ka, key=lambda e: -e.self_device_time_total)
for e in rows[:45]:
    print(f"{e.self_device_time_total / 1e3:9.2f} ms  n={e.count:6d}  {e.key[:110]}")

# ---- host pacing: CPU-side interval between consecutive forward launches (no GPU sync in the loop, so this is pure
# host time per lock-step round; it must stay below the GPU forward time for the GPU never to starve) ----
from eta_inversion_b200 import batching  # noqa: E402

stamps = []
_orig_run = batching.LockstepGroup._run


def _run(self):
    stamps.append(time.perf_counter())
    return _orig_run(self)


batching.LockstepGroup._run = _run
t0 = time.perf_counter(); step(); wall = 1e3 * (time.perf_counter() - t0)
import numpy as np  # noqa: E402
d = np.diff(np.array(stamps)) * 1e3
n = len(d)
print(f"wall {wall:.1f} ms, {len(stamps)} forwards; host interval between launches: inversion half mean {d[:n // 2].mean():.2f} ms "
      f"(p90 {np.percentile(d[:n // 2], 90):.2f}), edit half mean {d[n // 2:].mean():.2f} ms (p90 {np.percentile(d[n // 2:], 90):.2f}), "
      f"first launch at {1e3 * (stamps[0] - t0):.1f} ms, last launch at {1e3 * (stamps[-1] - t0):.1f} ms")
print("intervals (ms):", np.round(d[:12], 2), "...", np.round(d[n // 2 - 2:n // 2 + 10], 2), "...", np.round(d[-6:], 2))
