import sys, time, torch
sys.path.insert(0, '/root/repo')
import eta_inversion_b200 as etai
from eta_inversion_b200 import synthetic as syn
pipe, _ = etai.load_diffusion_model("synthetic-sd15", "cuda:0", variant="fp16")
img = syn.synthetic_image(0).cuda().to(pipe.vae.dtype)
lat = torch.randn(1, 4, 64, 64, device="cuda", dtype=pipe.vae.dtype)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / n
with torch.no_grad():
    print("encode 1:", t(lambda: pipe.vae.encode(img)))
    print("decode 1:", t(lambda: pipe.vae.decode(lat)))
    print("decode 2:", t(lambda: pipe.vae.decode(torch.cat([lat, lat]))))
    print("decode 8:", t(lambda: pipe.vae.decode(torch.cat([lat] * 8))))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        pipe.vae.decode(lat); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
