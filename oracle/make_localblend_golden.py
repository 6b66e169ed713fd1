"""ORACLE tooling (build container only): the reference's LocalBlend (modules/utils/ptp.py:18-47) run on a STRUCTURED
synthetic attention store -- per-word Gaussian blobs, so that max-pool -> nearest resize -> /max -> >0.3 -> OR of the
source and target rows gives a mixed mask -- and on a random latent pair.  The random-init UNet of the loop goldens has
spatially incoherent attention, so those goldens only ever see an all-ones LocalBlend mask; this fixture pins the masked
arithmetic itself.  Inputs are regenerated from the seed by `localblend_inputs` (shared with tests/test_localblend.py);
only the reference's outputs are stored (tests/golden/localblend.npz)."""
import os
import sys
import tempfile
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
PROMPTS = ["a cat sitting next to a mirror", "a tiger sitting next to a mirror"]
WORDS = [["cat"], ["tiger"]]
HEADS, LAYERS = 8, 5  # down_cross[2:4] + up_cross[:3] at 16x16 (ptp.py:33)


def localblend_inputs(seed: int = 0):
    """Per-layer cross-attention maps [LAYERS][prompts*HEADS, 256, 77] (prompt-major like the UNet batch) and x_t [2,4,64,64]."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(16.0), torch.arange(16.0), indexing="ij")
    layers = []
    for l in range(LAYERS):
        m = 0.02 * torch.rand((2 * HEADS, 256, 77), generator=g)
        for p in range(2):
            for h in range(HEADS):
                for w in range(1, 8):  # one blob per word token, jittered per (prompt, layer, head)
                    cx, cy = (3.0 + 1.7 * w + torch.rand(2, generator=g) * 1.5 + (2.0 if p else 0.0)).tolist()
                    sig = 1.2 + 0.4 * float(torch.rand(1, generator=g))
                    blob = torch.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * sig * sig)).reshape(256)
                    m[p * HEADS + h, :, w] += blob
        layers.append(m / m.sum(-1, keepdim=True))
    x_t = torch.randn((2, 4, 64, 64), generator=g)
    return layers, x_t


def main():
    os.system = lambda *a, **k: 0
    os.chdir(tempfile.mkdtemp(prefix="etai_oracle_"))
    sys.path[:0] = [str(REPO / "oracle" / "shim"), str(REPO), "/root/reference"]
    from modules.utils import ptp as ref_ptp
    from oracle.sd15 import SyntheticTokenizer
    model = SimpleNamespace(tokenizer=SyntheticTokenizer(), device=torch.device("cpu"),
                            scheduler=SimpleNamespace(num_inference_steps=10))
    out = {}
    for seed in (0, 1):
        layers, x_t = localblend_inputs(seed)
        lb = ref_ptp.LocalBlend(model, PROMPTS, WORDS)
        lb.counter = lb.start_blend  # the call below is the first one that blends
        store = {"down_cross": [None, None, layers[0], layers[1]], "up_cross": [layers[2], layers[3], layers[4]]}
        masks = []
        orig = ref_ptp.LocalBlend.get_mask

        def get_mask(self, x, maps, alpha, use_pool, _o=orig):
            m = _o(self, x, maps, alpha, use_pool)
            masks.append(m.clone())
            return m
        ref_ptp.LocalBlend.get_mask = get_mask
        y = lb(x_t.clone(), store)
        ref_ptp.LocalBlend.get_mask = orig
        out[f"x_out_{seed}"] = y.numpy()
        out[f"mask_{seed}"] = masks[0].numpy()
        print(f"seed {seed}: mask fractions (src, tgt rows) {masks[0].float().mean(dim=(1, 2, 3)).tolist()}")
    np.savez_compressed(REPO / "tests" / "golden" / "localblend.npz", **out)


if __name__ == "__main__":
    main()
