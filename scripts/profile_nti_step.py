"""One null-text-inversion inner step (train-mode forward + backward w.r.t. the context, B = 1, fp16) between
cudaProfilerStart / Stop, for `ncu --profile-from-start off --metrics gpu__time_duration.sum ...`; also prints the CUDA-event
time of the pair.    python scripts/profile_nti_step.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from eta_inversion_b200 import synthetic as syn  # noqa: E402
from eta_inversion_b200.engine import UNetEngine  # noqa: E402

eng = UNetEngine(syn.random_state_dict(syn.unet_param_spec(), 0), dtype=torch.float16, device="cuda:0", max_batch=4)
eng.enable_backward(1)
g = torch.Generator().manual_seed(0)
x = torch.randn((1, 4, 64, 64), generator=g).cuda()
ctx = torch.randn((1, 77, 768), generator=g).cuda()
w = (torch.randn((1, 4, 64, 64), generator=g) / 16384).cuda()
for _ in range(3):
    eng.forward_train(x, 501, ctx)
    eng.backward_ctx(w)
torch.cuda.synchronize()
a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
torch.cuda.cudart().cudaProfilerStart()
a.record()
eng.forward_train(x, 501, ctx)
b.record()
eng.backward_ctx(w)
c.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"train-mode forward {a.elapsed_time(b):.2f} ms, backward {b.elapsed_time(c):.2f} ms (B = 1, fp16, eager launches)")
