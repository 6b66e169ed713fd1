"""Micro-benchmarks of the exported ops through the C ABI (CUDA events, inputs rotated through > L2-size pools).
    python scripts/bench_ops.py [attn] [conv] [gemm] [gn] [ln] [--rows 16]
Prints achieved TFLOP/s or GB/s per shape against MEASURED_PEAKS.json."""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from eta_inversion_b200 import engine as E  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("what", nargs="*", default=["attn", "conv", "gemm", "gn", "ln"])
ap.add_argument("--rows", type=int, default=16)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--dtype", default="fp16")
args = ap.parse_args()
dt = {"fp16": torch.float16, "bf16": torch.bfloat16}[args.dtype]
B = args.rows
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
PT, PB = peaks.get("bf16_tflops", 1700.9), peaks.get("hbm_gbs", 6556.8)


def timeit(fn, n_variants):
    for i in range(3):
        fn(i % n_variants)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(args.iters):
        fn(i % n_variants)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / args.iters * 1e-3


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(dt)


if "attn" in args.what:
    for N, d in ((4096, 40), (1024, 80), (256, 160)):
        C = 8 * d
        nv = max(1, int(300e6 // (B * N * 3 * C * 2)) + 1)
        qkv = [rnd(B, N, 3 * C) for _ in range(nv)]
        t = timeit(lambda i: E.attention(qkv[i][:, :, :C], qkv[i][:, :, C:2 * C], qkv[i][:, :, 2 * C:], 8), nv)
        fl = 4.0 * N * N * C * B
        print(f"attention N={N} d={d} B={B}: {t * 1e6:8.1f} us  {fl / t / 1e12:7.1f} TFLOP/s ({fl / t / 1e12 / PT:.3f} of burst peak)")

if "conv" in args.what:
    for H, Ci, Co in ((64, 320, 320), (64, 640, 320), (32, 640, 640), (32, 1280, 640), (16, 1280, 1280), (16, 2560, 1280), (8, 1280, 1280), (8, 2560, 1280)):
        nv = max(1, int(300e6 // (B * H * H * Ci * 2)) + 1)
        xs = [rnd(B, H, H, Ci) for _ in range(nv)]
        w, bias = rnd(Co, 3, 3, Ci, scale=(9 * Ci) ** -0.5), rnd(Co)
        t = timeit(lambda i: E.conv3x3(xs[i], w, bias), nv)
        fl = 2.0 * B * H * H * Co * 9 * Ci
        print(f"conv3x3 {H}x{H} {Ci}->{Co} B={B}: {t * 1e6:8.1f} us  {fl / t / 1e12:7.1f} TFLOP/s ({fl / t / 1e12 / PT:.3f})")

if "gemm" in args.what:
    for HW, N, K, geglu in ((4096, 320, 320, 0), (4096, 960, 320, 0), (4096, 2560, 320, 1), (4096, 320, 1280, 0), (1024, 640, 640, 0),
                            (1024, 5120, 640, 1), (1024, 640, 2560, 0), (256, 1280, 1280, 0), (256, 10240, 1280, 1), (256, 1280, 5120, 0),
                            (64, 1280, 1280, 0)):
        M = B * HW
        nv = max(1, int(300e6 // (M * K * 2)) + 1)
        As = [rnd(M, K) for _ in range(nv)]
        W, bias = rnd(N, K, scale=K ** -0.5), rnd(N)
        t = timeit(lambda i: E.gemm(As[i], W, bias, geglu=bool(geglu)), nv)
        fl = 2.0 * M * N * K
        print(f"gemm M={M} N={N} K={K} geglu={geglu}: {t * 1e6:8.1f} us  {fl / t / 1e12:7.1f} TFLOP/s ({fl / t / 1e12 / PT:.3f})")

if "gn" in args.what:
    for HW, C in ((4096, 320), (4096, 960), (1024, 640), (1024, 1920), (256, 1280), (64, 2560)):
        nv = max(1, int(300e6 // (B * HW * C * 2)) + 1)
        xs = [rnd(B, HW, C) for _ in range(nv)]
        g, b_ = rnd(C), rnd(C)
        t = timeit(lambda i: E.groupnorm(xs[i], g, b_, 32, 1e-5, True), nv)
        by = 2.0 * B * HW * C * 2
        print(f"groupnorm+silu HW={HW} C={C} B={B}: {t * 1e6:8.1f} us  {by / t / 1e9:7.1f} GB/s algorithmic ({by / t / 1e9 / PB:.3f} of HBM peak)")

if "ln" in args.what:
    for HW, C in ((4096, 320), (1024, 640), (256, 1280)):
        nv = max(1, int(300e6 // (B * HW * C * 2)) + 1)
        xs = [rnd(B * HW, C) for _ in range(nv)]
        g, b_ = rnd(C), rnd(C)
        t = timeit(lambda i: E.layernorm(xs[i], g, b_), nv)
        by = 2.0 * B * HW * C * 2
        print(f"layernorm HW={HW} C={C} B={B}: {t * 1e6:8.1f} us  {by / t / 1e9:7.1f} GB/s algorithmic ({by / t / 1e9 / PB:.3f})")

if "sched" in args.what:
    # fused CFG + eta-DDIM step at a batched size: n latents of 4x64x64 fp32, eps rows [uncond.., cond..]
    for n in (16, 256, 4096):
        Eel = 4 * 64 * 64
        eps = torch.randn(2 * n, Eel, device="cuda")
        x = torch.randn(n, Eel, device="cuda")
        eta_map = (torch.rand(Eel, device="cuda") > 0.5).float()
        noise = torch.randn(1, Eel, device="cuda")
        t = timeit(lambda i: E.cfg_ddim_step(eps, x, 0.5, 0.6, 7.5, 0.3, 0.1, eta_map, noise, None), 1)
        by = (2 * n + n + n) * Eel * 4.0  # eps (2n rows) + x read, x' written
        print(f"cfg_ddim_step n={n} latents: {t * 1e6:8.1f} us  {by / t / 1e9:7.1f} GB/s algorithmic ({by / t / 1e9 / PB:.3f} of HBM peak)")
