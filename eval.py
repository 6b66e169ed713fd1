#!/usr/bin/env python
"""Sweep driver (BASELINE config 4): edits a set of (image, source prompt, target prompt) samples and writes one PNG
per sample, sharded PER IMAGE across the GPUs of one box.

    python eval.py --cfg cfg/eval/synthetic_pie.yaml                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 eval.py --cfg ...

Compared with the reference's eval.py (:27-195): the flags (--cfg --device --no_proc --override --skip_existing_dirs), the
YAML grid keys (data / model / method / edit_method / edit_cfg) and the result layout result/<cfg>/<nn_combo>/imgs/*.png are kept, finished PNGs are skipped unless --override
(eval_utils.py:260-263); the unit of parallelism is the sample (rank = index mod world, lock-step groups of --cobatch
inside a rank) instead of one process per config combination, and only per-sample records are gathered at the end.
The only dataset on this box is the synthetic PIE-shaped one (`data: [{type: synthetic_pie, n: 700}]`)."""
from __future__ import annotations

import argparse
import time
import itertools
import os
from pathlib import Path

import torch
import yaml

import eta_inversion_b200 as etai
from eta_inversion_b200 import synthetic as syn
from eta_inversion_b200.batching import run_lockstep, run_pipelined
from eta_inversion_b200.models import clone_pipeline
from eta_inversion_b200.sweep import run_sweep

PIE_DEFAULT_PTP = dict(is_replace_controller=False, cross_replace_steps={"default_": .4}, self_replace_steps=0.6)  # pie_bench_data.py:59-70
WORDS = ["cat", "tiger", "dog", "fox", "house", "castle", "car", "boat", "tree", "tower"]


def synthetic_pie_sample(i: int):
    """PIE-Bench-shaped record (keys of dataset/pie_bench_data.py:113-158) for sample i."""
    a, b = WORDS[i % len(WORDS)], WORDS[(i + 1 + i // len(WORDS)) % len(WORDS)]
    if a == b:
        b = WORDS[(WORDS.index(a) + 3) % len(WORDS)]
    src, tgt = f"a {a} sitting next to a mirror", f"a {b} sitting next to a mirror"
    cfg = {**PIE_DEFAULT_PTP, "prompts": [src, tgt], "blend_words": ((a,), (b,)), "equilizer_params": {"words": (b,), "values": (2,)}}
    return dict(name=f"{i:06d}", image=syn.synthetic_image(i), source_prompt=src, target_prompt=tgt, edit_cfg=cfg,
                edit_word_idx=(1, 1))


def combos(cfg):
    keys = ("model", "data", "method", "edit_method")
    for i, vals in enumerate(itertools.product(*[cfg[k] for k in keys])):
        yield i, dict(zip(keys, vals))


def respawn_per_device(devices, argv) -> int:
    """--device a b c (eval.py:163-175 of the reference starts one consumer per device): one rank per listed device under
    torch.distributed.run, samples sharded per image across them."""
    import subprocess
    import sys
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=",".join(devices))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={len(devices)}", "--master-addr",
           "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), sys.argv[0]] + argv
    return subprocess.call(cmd, env=env)


def main(cfg: str, cobatch: int, override: bool, prec: str, limit: int, pipes: int = 1, device=None, no_proc: bool = False,
         skip_existing_dirs: bool = False) -> None:
    import cv2
    if device and "RANK" not in os.environ:
        if len(device) > 1 and not no_proc:
            import sys
            argv, skip = [], 0
            for a in sys.argv[1:]:  # drop "--device a b c" from the child command line
                if a == "--device":
                    skip = 1
                    continue
                if skip and not a.startswith("--"):
                    continue
                skip = 0
                argv.append(a)
            raise SystemExit(respawn_per_device(device, argv))
        os.environ["CUDA_VISIBLE_DEVICES"] = device[0]  # --no_proc / one device: this process, first listed device
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec = yaml.safe_load(Path(cfg).read_text())
    root = Path("result") / Path(cfg).stem
    usd = syn.random_state_dict(syn.unet_param_spec(), 0)  # generated once, shared by the engine instances
    pipe, (_, postproc) = etai.load_diffusion_model("synthetic-sd15", f"cuda:{local}", variant=prec, max_batch=4 * cobatch,
                                                    unet_state_dict=usd)
    all_pipes = [pipe] + [clone_pipeline(pipe) for _ in range(max(1, pipes) - 1)]  # groups in flight per GPU
    del usd
    from concurrent.futures import ThreadPoolExecutor
    writer, writes = ThreadPoolExecutor(max_workers=4, thread_name_prefix="etai-png"), []  # PNG encode + write off the drivers
    for ci, combo in combos(spec):
        data = combo["data"] if isinstance(combo["data"], dict) else {"type": combo["data"]}
        n = min(int(data.get("n", 700)), limit) if limit else int(data.get("n", 700))
        out_dir = root / f"{ci:02d}_{combo['method']['type']}_{combo['edit_method']['type']}" / "imgs"
        if skip_existing_dirs:  # eval.py:41-44 of the reference: a combination whose directory exists is skipped whole
            exists = [out_dir.parent.exists()]
            if world > 1:
                dist.broadcast_object_list(exists, src=0)  # rank 0 decides before anybody creates the directory
                dist.barrier()
            if exists[0]:
                if rank == 0:
                    print(f"combo {ci}: {out_dir.parent} exists, skipped (--skip_existing_dirs)")
                continue
        out_dir.mkdir(parents=True, exist_ok=True)
        if rank == 0:
            (out_dir.parent / "cfg.yaml").write_text(yaml.safe_dump(combo))  # eval.py:46-47
        method, edit_method = dict(combo["method"]), dict(combo["edit_method"])

        def make_editor(p):
            inv = etai.load_inverter(model=p, **method)
            return etai.load_editor(inverter=inv, **edit_method)

        def run_groups(groups):
            """`pipes` lock-step groups in flight: one group's setup / VAE / host copies hide behind another's UNet."""
            samples = [[synthetic_pie_sample(i) for i in idx] for idx in groups]
            todo = [[s for s in ss if override or not (out_dir / f"{s['name']}.png").exists()] for ss in samples]
            jobs = [[dict(image=s["image"].cuda(), source_prompt=s["source_prompt"], target_prompt=s["target_prompt"],
                          cfg={**s["edit_cfg"]} if edit_method["type"] == "ptp" else None,
                          inv_cfg=dict(edit_word_idx=s["edit_word_idx"])) for s in td] for td in todo]
            live = [g for g, j in enumerate(jobs) if j]
            # one edit at a time: null-text inversion differentiates the UNet (its per-step optimisation has a data-dependent
            # length), so config 5 of BASELINE.json runs as independent replicas: one edit per GPU at a time, samples sharded
            # over the ranks.  (Plug-and-Play edits are co-batched like the others: the merged batch is laid out role-major.)
            if method.get("type") == "nti":
                with torch.no_grad():
                    res = {g: [make_editor(pipe).edit(**j) for j in jobs[g]] for g in live}
            elif len(all_pipes) > 1 and len(live) > 1:
                res = dict(zip(live, run_pipelined(all_pipes, [jobs[g] for g in live], make_editor)))
            else:
                res = {g: run_lockstep(pipe, jobs[g], make_editor) for g in live}
            out = []
            for g, ss in enumerate(samples):
                recs = {}
                for s, r in zip(todo[g], res.get(g, [])):
                    if r is None:
                        recs[s["name"]] = {"name": s["name"], "status": "unsupported"}
                        continue
                    # device -> host uint8 here (orders after the group's stream); PNG encode + file write on the writer pool
                    # (cv2 releases the GIL), so the pipe's driver thread goes straight on to its next group
                    img = postproc(r["image"])
                    writes.append(writer.submit(lambda path=str(out_dir / f"{s['name']}.png"), im=img:
                                                cv2.imwrite(path, cv2.cvtColor(im, cv2.COLOR_RGB2BGR))))
                    lat = r["latent"][0] if isinstance(r["latent"], (list, tuple)) else r["latent"]  # EDICT: coupled pair
                    recs[s["name"]] = {"name": s["name"], "status": "done", "latent_mean": float(lat.mean())}
                out.append([recs.get(s["name"], {"name": s["name"], "status": "skipped"}) for s in ss])
            return out

        def run_group(idx):
            return run_groups([idx])[0]
        if world > 1:
            dist.barrier()
        t_sweep = time.perf_counter()
        records = run_sweep(n, rank, world, cobatch, run_group, window=4 * len(all_pipes), run_groups=run_groups)
        for w in writes:  # every PNG of this rank is on disk before the sweep counts as finished
            if not w.result():
                raise RuntimeError("cv2.imwrite failed")
        writes.clear()
        if world > 1:
            dist.barrier()
        t_sweep = time.perf_counter() - t_sweep  # includes the final gather and the PNG writes of the slowest rank
        if rank == 0:
            done = sum(r['status'] == 'done' for r in records)
            print(f"combo {ci}: {done} edits in {t_sweep:.1f} s on {world} GPU(s) = {done / max(t_sweep, 1e-9):.2f} edits/s "
                  f"(first-call graph capture, VAE / CLIP, PNG encode + write included; model load excluded)")
            (out_dir.parent / "records.yaml").write_text(yaml.safe_dump(records))
            print(f"combo {ci}: {sum(r['status'] == 'done' for r in records)} edited, "
                  f"{sum(r['status'] == 'skipped' for r in records)} skipped -> {out_dir}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Run an editing sweep sharded per image across the visible GPUs.")
    ap.add_argument("--cfg", required=True, help="Config file for evaluation.")
    ap.add_argument("--cobatch", type=int, default=4, help="Edits walked in lock step per GPU.")
    ap.add_argument("--device", nargs="+", help="Which cuda devices to use. Can be multiple (one rank per device).")
    ap.add_argument("--no_proc", action="store_true", help="Disables multiprocessing (runs in this process, first device).")
    ap.add_argument("--override", action="store_true", help="Override old results.")
    ap.add_argument("--skip_existing_dirs", action="store_true", help="Skips existing directories.")
    ap.add_argument("--prec", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--limit", type=int, default=0, help="Only the first N samples (smoke runs).")
    ap.add_argument("--pipes", type=int, default=3, help="Lock-step groups in flight per GPU (each on its own engine handle).")
    main(**vars(ap.parse_args()))
