"""LocalBlend (reference modules/utils/ptp.py:18-47) on a structured attention store: the engine's LocalBlend consumes the
head/layer-summed store the attention kernels accumulate; the golden was written by the reference's class from the
per-layer maps (oracle/make_localblend_golden.py).  Mixed masks (18 % / 27 % of the pixels), unlike the loop goldens whose
random-init attention gives all-ones masks."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.make_localblend_golden import HEADS, PROMPTS, WORDS, localblend_inputs

pytestmark = pytest.mark.gpu


class _Store:
    def __init__(self, acc):
        self.acc = acc

    def accumulated(self, res, from_where):
        assert res == 16 and tuple(from_where) == ("down", "up")
        return self.acc


@pytest.mark.parametrize("seed", [0, 1])
def test_localblend_matches_reference_on_structured_maps(seed):
    from eta_inversion_b200.models import SyntheticTokenizer
    from eta_inversion_b200.utils.ptp import LocalBlend
    gold = np.load(GOLDEN / "localblend.npz")
    layers, x_t = localblend_inputs(seed)
    dev = torch.device("cuda:0")
    model = SimpleNamespace(tokenizer=SyntheticTokenizer(), device=dev, scheduler=SimpleNamespace(num_inference_steps=10))
    lb = LocalBlend(model, PROMPTS, WORDS)
    lb.counter = lb.start_blend
    # what the fused cross-attention store holds: sum over the 5 layers x 8 heads, per prompt row
    acc = sum(m.reshape(2, HEADS, 256, 77).sum(1) for m in layers).to(dev)
    masks = []
    orig = lb.get_mask
    lb.get_mask = lambda x, maps, use_pool=True: masks.append(orig(x, maps, use_pool)) or masks[-1]
    y = lb(x_t.to(dev), _Store(acc)).cpu()
    mask, gmask = masks[0].cpu(), torch.from_numpy(gold[f"mask_{seed}"])
    frac = mask.float().mean(dim=(1, 2, 3)).tolist()
    agree = (mask == gmask).float().mean().item()
    print(f"seed {seed}: mask fractions {frac}, agreement with the reference mask {agree:.5f}")
    assert 0.05 < min(frac) and max(frac) < 0.95      # a genuinely mixed mask
    assert agree >= 0.999                              # summation order may flip a pixel exactly at the threshold
    same = (mask == gmask).all(dim=0, keepdim=True).expand_as(y[:, :1]).expand_as(y)
    assert torch.equal(y[same], torch.from_numpy(gold[f"x_out_{seed}"])[same])
