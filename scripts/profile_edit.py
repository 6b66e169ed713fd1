"""Profiling driver: one etainv+ptp edit with `--steps` DDIM steps between cudaProfilerStart/Stop (after a warm-up edit).

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python scripts/profile_edit.py --steps 1
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc -c 3 \
      -o gpurun_out/prof python scripts/profile_edit.py --steps 1
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import eta_inversion_b200 as etai  # noqa: E402
from eta_inversion_b200 import synthetic as syn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--variant", default="fp16")
ap.add_argument("--cobatch", type=int, default=1, help="edits sharing each UNet forward (bench.py default: 4)")
args = ap.parse_args()

cfg = dict(is_replace_controller=True, cross_replace_steps={"default_": .8}, self_replace_steps=.5,
           blend_words=[["cat"], ["tiger"]], equilizer_params={"words": ["tiger"], "values": [2]})
pipe, _ = etai.load_diffusion_model("synthetic-sd15", "cuda:0", variant=args.variant, max_batch=4 * args.cobatch)


def make_editor(lane_pipe):
    inv = etai.load_inverter(type="etainv", model=lane_pipe, scheduler="ddim", num_inference_steps=args.steps)
    return etai.load_editor(type="ptp", inverter=inv)


ed = make_editor(pipe)
img = syn.synthetic_image(0).cuda()


def edit():
    if args.cobatch > 1:
        from eta_inversion_b200.batching import run_lockstep
        job = dict(image=img, source_prompt="a cat sitting next to a mirror", target_prompt="a tiger sitting next to a mirror",
                   cfg={**cfg}, inv_cfg=dict(edit_word_idx=(1, 1)))
        return run_lockstep(pipe, [dict(job, cfg={**cfg}) for _ in range(args.cobatch)], make_editor)
    with torch.no_grad():
        return ed.edit(img, "a cat sitting next to a mirror", "a tiger sitting next to a mirror", cfg={**cfg},
                       inv_cfg=dict(edit_word_idx=(1, 1)))


edit()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
edit()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one edit of", args.steps, "step(s)")
