"""Build-container experiment (needs /root/reference): the reference's OWN loops with the oracle model cast to bf16 / fp16
on the CPU, against the fp32 golden image of the same scenario.  Tells what PSNR the 16-bit arithmetic itself allows on
the synthetic model, i.e. what the engine's 35 dB gate can be compared with.
    python scripts/ref_16bit_psnr.py --scenario etainv_ptp_refine_3 --dtype bf16"""
import argparse
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from oracle import run_reference as rr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", nargs="+", default=["etainv_ptp_refine_3"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--json", default=None, help="merge the results into this JSON file (tests/golden/ref_16bit_psnr.json)")
    args = ap.parse_args()
    rr._setup_paths()
    torch.set_num_threads(os.cpu_count())
    from eta_inversion_b200 import synthetic as syn
    from eta_inversion_b200.metrics import psnr
    from oracle import sd15
    import modules
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[args.dtype]
    pipe = sd15.build_pipeline(syn.random_state_dict(syn.unet_param_spec(), 0), syn.random_state_dict(syn.vae_param_spec(), 1),
                               seed=0, dtype=dt)
    pipe.text_encoder.to(dt)
    import json
    results = {}
    for name in args.scenario:
        inv_kw, ed_type, ed_kw, cfg, inv_cfg = rr.SCENARIOS[name]
        inverter = modules.load_inverter(model=pipe, **inv_kw)
        editor = modules.load_editor(inverter=inverter, type=ed_type, **ed_kw)
        t0 = time.perf_counter()
        with torch.no_grad():
            res = editor.edit(syn.synthetic_image(0), rr.SRC, rr.TGT, cfg=None if cfg is None else {**cfg}, inv_cfg=inv_cfg)
        gold = np.load(REPO / "tests" / "golden" / f"{name}.npz")
        p = psnr(res["image"].float(), torch.from_numpy(gold["image_f16"]).float())
        results[name] = round(float(p), 2)
        print(f"{name}: reference loops + oracle model in {args.dtype} on CPU vs fp32 golden image: PSNR {p:.1f} dB "
              f"({time.perf_counter() - t0:.0f} s)", flush=True)
    if args.json:
        path = Path(args.json) if os.path.isabs(args.json) else REPO / args.json
        data = json.loads(path.read_text()) if path.exists() else {}
        data.setdefault(args.dtype, {}).update(results)
        path.write_text(json.dumps(data, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
