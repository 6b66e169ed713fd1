"""ORACLE tooling (build container only): run the reference's own inverse schedulers on small seeded tensors and write
tests/golden/schedulers.npz.  Reference: modules/inverse_schedulers/ddpm_inverse_scheduler.py:9-203,
modules/inversion/edict_inversion.py:17-222 (over the restated diffusers DDIMScheduler of oracle/sd15.py)."""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
os.system = lambda *a, **k: 0
os.chdir(tempfile.mkdtemp(prefix="etai_oracle_"))
sys.path[:0] = [str(REPO / "oracle" / "shim"), str(REPO), "/root/reference"]

from diffusers import DDIMScheduler  # noqa: E402  (the shim)
from modules.inverse_schedulers import DDPMInverseScheduler  # noqa: E402
from modules.inversion.edict_inversion import EdictScheduler, EdictSchedulerInverse  # noqa: E402

SD = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000, clip_sample=False,
          set_alpha_to_one=False, steps_offset=1)
out = {}
g = torch.Generator().manual_seed(11)
z0 = torch.randn((1, 4, 8, 8), generator=g)
eps = torch.randn((6, 1, 4, 8, 8), generator=g)
out["z0"], out["eps"] = z0.numpy(), eps.numpy()
for steps in (6, 3):          # 3: 1000/3 is fractional -> EDICT interpolates alpha_bar
    for markov in (False, True):
        base = DDIMScheduler(**SD)
        base.set_timesteps(steps)
        inv = DDPMInverseScheduler.from_scheduler(base, markovian_forward=markov)
        inv.set_timesteps(steps)
        xts = inv.sample_latents(z0, generator=torch.Generator().manual_seed(5))
        key = f"ddpm_s{steps}_m{int(markov)}"
        out[key + "_xts"] = xts.numpy()
        zs, xs = [], []
        for i, t in enumerate(inv.timesteps):
            r = inv.step(eps[i % 6], t, inv.get_sampled_latent_by_t(xts, t), xts)
            zs.append(r.variance_noise.numpy())
            xs.append(r.prev_sample.numpy())
        out[key + "_z"], out[key + "_x"] = np.stack(zs), np.stack(xs)
        out[key + "_var"] = np.array([float(inv.get_variance(int(t))) for t in inv.timesteps])
    base = DDIMScheduler(**SD)
    base.set_timesteps(steps)
    bwd, fwd = EdictScheduler(base), EdictSchedulerInverse(base)
    x = z0.clone()
    f_out, b_out = [], []
    for i, t in enumerate(fwd.timesteps):
        x = fwd.step(eps[i % 6], t, x).prev_sample
        f_out.append(x.numpy())
    for i, t in enumerate(bwd.timesteps):
        x = bwd.step(eps[(steps - 1 - i) % 6], t, x).prev_sample
        b_out.append(x.numpy())
    out[f"edict_s{steps}_fwd"], out[f"edict_s{steps}_bwd"] = np.stack(f_out), np.stack(b_out)
np.savez_compressed(REPO / "tests" / "golden" / "schedulers.npz", **out)
print("wrote schedulers.npz", {k: v.shape for k, v in out.items()})
