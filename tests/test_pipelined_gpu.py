"""run_pipelined: two lock-step groups in flight on two engines / streams give the same results as running the groups
one after the other on one engine."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SRC, TGT = "a cat sitting next to a mirror", "a tiger sitting next to a mirror"
CFG = dict(is_replace_controller=True, cross_replace_steps={"default_": .8}, self_replace_steps=.5,
           blend_words=[["cat"], ["tiger"]], equilizer_params={"words": ["tiger"], "values": [2]})


def test_pipelined_groups_match_sequential():
    import eta_inversion_b200 as etai
    from eta_inversion_b200 import synthetic as syn
    from eta_inversion_b200.batching import run_lockstep, run_pipelined
    from eta_inversion_b200.models import clone_pipeline

    pipe, _ = etai.load_diffusion_model("synthetic-sd15", "cuda:0", variant="fp32", max_batch=8)
    # VAE and text tower are native and bit-reproducible (fixed-order reductions), like the UNet: no memoisation needed
    pipe2 = clone_pipeline(pipe)

    def make_editor(p):
        inv = etai.load_inverter(type="etainv", model=p, scheduler="ddim", num_inference_steps=3)
        return etai.load_editor(type="ptp", inverter=inv)

    def jobs(i0):
        return [dict(image=syn.synthetic_image(i0 + i).cuda(), source_prompt=SRC, target_prompt=TGT, cfg={**CFG},
                     inv_cfg=dict(edit_word_idx=(1, 1))) for i in range(2)]

    groups = [jobs(0), jobs(2), jobs(4)]
    seq = [run_lockstep(pipe, [dict(j, cfg={**CFG}) for j in g], make_editor) for g in groups]
    torch.cuda.synchronize()
    par = run_pipelined([pipe, pipe2], [[dict(j, cfg={**CFG}) for j in g] for g in groups], make_editor)
    torch.cuda.synchronize()
    for gi, (gs, gp) in enumerate(zip(seq, par)):
        for li, (a, b) in enumerate(zip(gs, gp)):
            print(f"group {gi} lane {li}: latent diff {(a['latent'] - b['latent']).abs().max().item():.3e} "
                  f"latent_inv diff {(a['latent_inv'] - b['latent_inv']).abs().max().item():.3e}")
    for gs, gp in zip(seq, par):
        for a, b in zip(gs, gp):
            assert torch.equal(a["latent"], b["latent"])       # fp32 path: bit-exact, whichever engine / stream ran it
            # the torch VAE decode (cuDNN) is not bit-reproducible across streams; the engine path above is
            assert (a["image"] - b["image"]).abs().max().item() < 1e-4
