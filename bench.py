#!/usr/bin/env python
"""Benchmark of the hot path: edits/sec for etainv + prompt-to-prompt (word-swap replace), SD-1.5 architecture,
512x512, 50 DDIM steps, CFG 7.5, fp16 (BASELINE.json metric / configs[1]) on N B200 GPUs of one node.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference's CPU path (oracle port) on the host cores

One "step" = one complete `editor.edit(...)` of one (image, prompt pair): CLIP x4, VAE encode, 50-step eta inversion
(B=2), 50-step PtP edit loop (B=4), VAE decode x2 -- exactly what edit_image.py:113-115 of the reference times.
Ranks process different synthetic pairs and never communicate inside the loop (weak scaling); value = N*K / time.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SRC, TGT = "a cat sitting next to a mirror", "a tiger sitting next to a mirror"
PTP_CFG = dict(is_replace_controller=True, cross_replace_steps={"default_": .8}, self_replace_steps=.5,
               blend_words=[["cat"], ["tiger"]], equilizer_params={"words": ["tiger"], "values": [2]})
INV_CFG = dict(edit_word_idx=(1, 1))
ROWS_PER_EDIT = lambda steps: steps * 2 + steps * 4  # noqa: E731  (eta_inversion.py:319-328: B=2 inversion, B=4 edit)
# rows the engine computes: at guidance_scale_fwd = 1 (the default) the inversion's unconditional row has weight exactly 0
# in the reference's combine u + 1*(c - u) and is not computed (EtaInversion.skip_zero_weight_uncond)
ROWS_COMPUTED = lambda steps, skip: steps * (1 if skip else 2) + steps * 4  # noqa: E731


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1395.7), d.get("hbm_gbs", 6556.8), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        util = [float(r[7]) for r in self.rows if len(r) >= 8 and r[7].replace(".", "").isdigit()]
        power = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "gpu_util_pct_mean": round(statistics.mean(util), 1) if util else None,
                "power_w_mean": round(statistics.mean(power), 1) if power else None}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle restatement of the reference loop on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference(inv_steps: int, steps: int, warmup: int, budget_s: float = 240.0):
    """Times the reference's CPU path (oracle/ref_loop.py over oracle/sd15.py, fp32, all host threads) on a BOUNDED
    sample per step: one DDIM step of each loop (1 UNet call at B=2 with the attention store + 1 at B=4 with the PtP
    hooks, attention materialised like the reference) = 1/inv_steps of an edit's UNet work; VAE/CLIP measured once.
    edits/s = 1 / (fixed + inv_steps * per_step)."""
    from eta_inversion_b200 import synthetic as syn
    from oracle import ref_loop, sd15
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    pipe = sd15.build_pipeline(syn.random_state_dict(syn.unet_param_spec(), 0),
                               syn.random_state_dict(syn.vae_param_spec(), 1), seed=0)
    img = syn.synthetic_image(0)
    unet_s = [0.0]

    def on_unet(orig, *a, **k):
        t0 = time.perf_counter()
        out = orig(*a, **k)
        unet_s[0] += time.perf_counter() - t0
        return out
    per_step, fixed, t_start = [], None, time.perf_counter()
    for i in range(warmup + steps):
        if time.perf_counter() - t_start > budget_s and len(per_step) >= 1:
            break
        unet_s[0] = 0.0
        t0 = time.perf_counter()
        ref_loop.edit(pipe, img, SRC, TGT, inverter="etainv", editor="ptp", steps=1, ptp_cfg=PTP_CFG,
                      decode=(fixed is None), on_unet=on_unet)
        total = time.perf_counter() - t0
        if fixed is None:
            fixed = total - unet_s[0]  # CLIP x4 + VAE encode + 2 decodes + scheduler algebra, measured once
        if i >= min(warmup, 1):  # CPU has no lazy init worth 3 warm-ups of ~10 s each: 1 untimed step, rest timed
            per_step.append(unet_s[0])
    step_s = statistics.mean(per_step)
    edit_s = fixed + inv_steps * step_s
    return {"value": 1.0 / edit_s, "edit_seconds": edit_s, "per_ddim_step_pair_s": step_s, "fixed_s": fixed,
            "cores": cores, "timed_samples": len(per_step)}


def cpu_reference_whole_edit(inv_steps: int, warmup: int):
    """The reference arm proper: ONE complete edit (CLIP x4, VAE encode, `inv_steps` eta-inversion steps at B=2,
    `inv_steps` PtP steps at B=4 with the attention materialised like the reference, 2 VAE decodes) timed with
    time.perf_counter around the edit call, exactly the region edit_image.py:113-115 of the reference times; fp32, all
    host threads.  `warmup` untimed samples of one DDIM step of each loop come first (thread pools, page-in)."""
    from eta_inversion_b200 import synthetic as syn
    from oracle import ref_loop, sd15
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    pipe = sd15.build_pipeline(syn.random_state_dict(syn.unet_param_spec(), 0),
                               syn.random_state_dict(syn.vae_param_spec(), 1), seed=0)
    img = syn.synthetic_image(0)
    for _ in range(warmup):
        ref_loop.edit(pipe, img, SRC, TGT, inverter="etainv", editor="ptp", steps=1, ptp_cfg=PTP_CFG, decode=False)
    t0 = time.perf_counter()
    ref_loop.edit(pipe, img, SRC, TGT, inverter="etainv", editor="ptp", steps=inv_steps, ptp_cfg=PTP_CFG, decode=True)
    return time.perf_counter() - t0, cores


def bench_config(inv_steps):
    """`config` is identical in both arms (the driver compares them)."""
    return {"workload": workload_name(inv_steps), "inv_steps": inv_steps, "unet_rows_per_edit": ROWS_PER_EDIT(inv_steps),
            "l2": "GPU arm: no explicit flush, every UNet forward streams 1.72 GB of fp16 weights (>> 126 MB L2); "
                  "CPU arm: not applicable"}


def run_reference(args, rank):
    """`--impl reference`: K steps, each 1/K of ONE whole edit that is timed from start to end (no extrapolation):
    ms_per_step * K = the edit's wall time, value = 1 / that."""
    if rank != 0:
        return
    K, W = max(args.steps, 1), max(args.warmup, 0)
    edit_s, cores = cpu_reference_whole_edit(args.inv_steps, W)
    sample = (f"one whole {args.inv_steps}+{args.inv_steps}-step etainv+PtP edit (CLIP x4, VAE encode, inversion B=2, edit B=4 "
              f"with materialised attention, 2 VAE decodes), fp32, torch on {cores} host threads, wall clock around the edit "
              f"call; reported as {K} steps of 1/{K} edit each")
    line = {"impl": "reference", "metric": "edits/sec etainv+PtP SD1.5 512^2 50-step", "value": 1.0 / edit_s, "unit": "edits/s",
            "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1000.0 * edit_s / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": bench_config(args.inv_steps),
            "edit_seconds": edit_s,
            "cpu_baseline": {"value": 1.0 / edit_s, "unit": "edits/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": 1.0 / edit_s, "unit": "edits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(inv_steps):
    return (f"etainv + ptp (word-swap replace, cross .8 / self .5, LocalBlend, reweight x2), SD-1.5 architecture "
            f"random-init, 512x512 synthetic image, {inv_steps} DDIM steps, CFG 7.5 (BASELINE configs[1])")


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="etai", choices=["etai", "reference"])
    ap.add_argument("--inv-steps", type=int, default=50, dest="inv_steps")
    ap.add_argument("--variant", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cobatch", type=int, default=4,
                    help="independent edits walked in lock step per step on each GPU (they share every UNet forward)")
    ap.add_argument("--reference-rows", action="store_true", dest="reference_rows",
                    help="also compute the inversion's zero-weight unconditional rows, like the reference (300 instead of 250 "
                         "UNet rows per edit; same results)")
    ap.add_argument("--pipes", type=int, default=3,
                    help="lock-step groups in flight per GPU, each on its own engine (activation arena, graphs, stream): one "
                         "group's setup / VAE / host copies overlap the other's UNet forwards")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch.distributed as dist
    import eta_inversion_b200 as etai
    from eta_inversion_b200 import engine as E, flops, synthetic as syn
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(args.warmup, 3)
    K = args.steps

    from eta_inversion_b200.batching import run_lockstep, run_pipelined
    from eta_inversion_b200.inversion.eta_inversion import EtaInversion
    from eta_inversion_b200.models import clone_pipeline
    EtaInversion.skip_zero_weight_uncond = not args.reference_rows
    CB = max(1, args.cobatch)
    G = max(1, args.pipes)  # a step = G groups of CB edits, the groups of all steps are pipelined over G engines
    usd = syn.random_state_dict(syn.unet_param_spec(), 0)  # generated once, shared by the G engine instances
    pipe, (preproc, postproc) = etai.load_diffusion_model("synthetic-sd15", f"cuda:{local}", variant=args.variant,
                                                          max_batch=4 * CB, unet_state_dict=usd)
    pipes = [pipe] + [clone_pipeline(pipe) for _ in range(G - 1)]
    del usd
    for p_ in pipes:
        p_.cache_text_embeddings = False  # the reference's timed region contains the 4 CLIP passes of every edit

    def make_editor(p):
        inv = etai.load_inverter(type="etainv", model=p, scheduler="ddim", num_inference_steps=args.inv_steps,
                                 guidance_scale_bwd=7.5)
        return etai.load_editor(type="ptp", inverter=inv)
    editor = make_editor(pipe)
    n_img = (W + K) * CB * G
    host_imgs = [syn.synthetic_image(100000 * rank + i).pin_memory() for i in range(n_img)]
    dev_imgs = [h.cuda(non_blocking=True) for h in host_imgs]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def edit(imgs):
        """One lock-step group = CB independent edits (different images) sharing every UNet forward."""
        if CB == 1:
            with torch.no_grad():
                return [editor.edit(imgs[0], SRC, TGT, cfg={**PTP_CFG}, inv_cfg=INV_CFG)]
        jobs = [dict(image=im, source_prompt=SRC, target_prompt=TGT, cfg={**PTP_CFG}, inv_cfg=dict(INV_CFG)) for im in imgs]
        return run_lockstep(pipe, jobs, make_editor)

    def batch(lst, i):
        return lst[i * CB:(i + 1) * CB]

    def steps(imgs, first, count, on_result=None):
        """`count` steps starting at step `first`: count * G groups, G of them in flight at any time."""
        groups = [batch(imgs, g) for g in range(first * G, (first + count) * G)]
        if G == 1:
            out = []
            for gi, grp in enumerate(groups):
                r = edit(grp)
                out.append(on_result(gi, r) if on_result else r)
            return out
        jobs = [[dict(image=im, source_prompt=SRC, target_prompt=TGT, cfg={**PTP_CFG}, inv_cfg=dict(INV_CFG)) for im in grp]
                for grp in groups]
        return run_pipelined(pipes, jobs, make_editor, on_result=on_result)

    def launches_now():  # UNet handles + the shared VAE / CLIP handles + the op-level entry points (scheduler step, noise losses)
        return (sum(p.unet.launch_count for p in pipes) + pipe.vae.launch_count + pipe.text_encoder.launch_count + E.LAUNCHES[0])

    # ---- (1) device-resident throughput ----------------------------------------------------------
    steps(dev_imgs, 0, W)
    barrier()
    l0 = launches_now()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record()
        steps(dev_imgs, W, K)       # returns with every group's stream synchronised
        torch.cuda.synchronize()
        ev1.record()
        barrier()
    launches = launches_now() - l0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())

    # ---- (2) end to end through the public API with host buffers -----------------------------------
    d2h_box = [0]

    def to_host(gi, results):
        n = 0
        for res in results:
            out = [postproc(res["image"]), postproc(res["image_inv"])]  # device -> host uint8 images (cv2.imwrite input)
            n += sum(o.nbytes for o in out)  # the bytes that actually cross PCIe: HWC uint8 (converted on the device)
        d2h_box[0] = n * G  # per step
        return None

    barrier()
    t0 = time.perf_counter()
    steps(host_imgs, W, K, on_result=to_host)   # pinned host images: every lane uploads its own inside the region
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    d2h = d2h_box[0]
    h2d = host_imgs[0].numel() * host_imgs[0].element_size() * CB * G

    # ---- (2b) share of the step spent inside UNet forwards (graph replay, CUDA events), one more step ----
    pipe.unet.time_forwards(True)
    barrier()
    t0 = time.perf_counter()
    edit(batch(dev_imgs, W))
    barrier()
    step_wall_ms = (time.perf_counter() - t0) * 1e3
    unet_graph_ms = pipe.unet.time_forwards(False)

    # ---- (3) per-category device time of one edit (CUDA events around every op; not part of the timings above) ---
    pipe.unet.profile(True)
    edit(batch(dev_imgs, W))
    torch.cuda.synchronize()
    prof = pipe.unet.profile(False)
    for v in prof.values():
        v["ms"] /= CB  # per edit

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    tf_peak, hbm_peak, peak_src = peaks()
    macs = flops.unet_macs_per_row()
    rows = ROWS_COMPUTED(args.inv_steps, not args.reference_rows)
    unet_ms = sum(v["ms"] for v in prof.values())
    # conv3x3 and the dense projections are ONE kernel (gemm_tc_k<T,BN,CONV>: the conv instantiation only differs in how the
    # TMA producer addresses the A operand), so they are accounted together as the dominant kernel
    gemm_ms = prof["conv3x3"]["ms"] + prof["gemm"]["ms"]
    if gemm_ms >= prof["self_attn"]["ms"]:
        dom, dom_ms, dom_tflop = "conv3x3", gemm_ms, 2.0 * (macs["conv3x3"] + macs["gemm"]) * rows / 1e12
        dom_name = "gemm_tc_k (implicit-GEMM conv3x3 + dense instantiations)"
    else:
        dom, dom_ms, dom_tflop = "self_attn", prof["self_attn"]["ms"], 2.0 * macs["self_attn"] * rows / 1e12
        dom_name = "attn_d40_k / attn_tc2_k / attn_tc_k"
    achieved = dom_tflop / (dom_ms / 1e3)
    breakdown = {k: {"ms_per_edit": round(v["ms"], 2), "launches": v["launches"],
                     "tflops": round(2.0 * macs[k] * rows / 1e12 / (v["ms"] / 1e3), 1) if k in macs and v["ms"] > 0 else None}
                 for k, v in prof.items()}
    # DRAM traffic of the dominant kernel: one `ncu --set full` capture per round, kept under profiles/ (per launch)
    traffic_bytes, traffic_note = None, None
    tpath = Path(__file__).resolve().parent / "profiles" / "r02_ncu_conv_traffic.json"
    if dom == "conv3x3" and tpath.exists():
        tj = json.loads(tpath.read_text())
        traffic_bytes = int(tj["dram_bytes_read"] + tj["dram_bytes_write"])
        traffic_note = (f"{tj['launch']}: {traffic_bytes / 1e6:.1f} MB DRAM traffic vs {tj['algorithmic_bytes'] / 1e6:.1f} MB "
                        f"algorithmic (in + weights + out) in {tj['duration_us']} us under ncu; {tj['note']}")
    value = world * K * CB * G / (ms_total / 1e3)
    line = {
        "metric": "edits/sec etainv+PtP SD1.5 512^2 50-step", "value": value, "unit": "edits/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.variant, "data": "synthetic",
        "config": bench_config(args.inv_steps),
        "engine": {"edits_per_step": CB * G,
                   "cobatch": f"{CB} independent edits per GPU share each UNet forward (B={(2 if args.reference_rows else 1) * CB} "
                              f"inversion, B={4 * CB} edit)",
                   "pipes": f"{G} lock-step group(s) in flight per GPU, each on its own engine handle (one shared copy of the weights)",
                   "unet_rows_computed_per_edit": rows,
                   "rows_note": ("the reference's etainv inversion always runs [uncond, cond] although guidance_scale_fwd = 1 gives the "
                                 "unconditional row weight 0 (eta_inversion.py:319-328); the engine does not compute that row "
                                 "(identical result; --reference-rows computes it)") if not args.reference_rows else
                                "all rows the reference computes",
                   "tflop_per_edit": round(2.0 * macs["total"] * rows / 1e12, 2),
                   "ms_per_unet_forward_avg": round(unet_graph_ms / (2 * args.inv_steps), 3),
                   "unet_share_of_step": round(unet_graph_ms / step_wall_ms, 3),
                   "unet_share_note": "UNet-forward share of the wall time of ONE lock-step group running ALONE (nothing overlaps its "
                                      "host work; depends on the box's host speed); in the timed regions the groups in flight hide "
                                      "that time: see clocks.gpu_util_pct_mean",
                   "parallelism": f"per-image sharding, {world} independent rank(s), no collective in the loop"},
        "clocks": clocks.summary(),
        "e2e": {"value": world * K * CB * G / e2e_s, "unit": "edits/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": dom_name,
                     "achieved": round(achieved, 1), "peak": tf_peak, "unit": "TFLOP/s", "frac": round(achieved / tf_peak, 4),
                     "traffic": traffic_bytes, "traffic_note": traffic_note, "peak_source": peak_src,
                     "note": f"{dom_tflop:.1f} TFLOP per edit ({rows} UNet rows) / {dom_ms:.1f} ms of CUDA-event time per edit",
                     "unet_tflops_all_kernels": round(2.0 * macs["total"] * rows / 1e12 / (unet_ms / 1e3), 1),
                     "breakdown": breakdown},
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference(args.inv_steps, steps=1, warmup=1, budget_s=60.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": "edits/s", "cores": r["cores"], "kind": "port",
                                    "sample": f"{r['timed_samples']} sample(s) of 1/{args.inv_steps} edit (1 inversion + 1 PtP edit "
                                              f"UNet step, attention materialised) + VAE/CLIP once, extrapolated; "
                                              f"{r['edit_seconds']:.0f} s per edit"}
        except Exception as e:  # the baseline is a report, never a reason to lose the measurement
            line["cpu_baseline"] = {"value": None, "unit": "edits/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {type(e).__name__}: {e}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
