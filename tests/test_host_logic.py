"""CPU tests of the host-side logic (no GPU, no CUDA calls): token alignment / alpha tables / eta schedule against
goldens produced by the reference's own functions (oracle/make_host_goldens.py), scheduler constants against the
oracle, the PSNR known answer of the reference's metric test, sharding, and the analytic FLOP count."""
import json
import math

import numpy as np
import pytest
import torch

from conftest import GOLDEN

H = json.loads((GOLDEN / "host_logic.json").read_text())


@pytest.fixture(scope="module")
def tok():
    from eta_inversion_b200.models import SyntheticTokenizer
    return SyntheticTokenizer()


def test_tokenizer_matches_oracle_tokenizer(tok):
    from oracle.sd15 import SyntheticTokenizer as OracleTok
    o = OracleTok()
    for s, t in H["pairs"]:
        for text in (s, t, ""):
            assert tok.encode(text) == o.encode(text)
            assert torch.equal(tok([text]).input_ids, o([text]).input_ids)
    assert tok("a b").input_ids.shape == (1, 77)


def test_refinement_and_replacement_mappers_match_reference(tok):
    from eta_inversion_b200.utils import seq_aligner
    for (s, t), ref, rep in zip(H["pairs"], H["refine"], H["replace"]):
        m, a = seq_aligner.get_refinement_mapper([s, t], tok)
        assert m[0].tolist() == ref["mapper"] and a[0].tolist() == ref["alphas"]
        if rep is not None:
            assert seq_aligner.get_replacement_mapper([s, t], tok)[0].tolist() == rep
        else:
            with pytest.raises(ValueError):
                seq_aligner.get_replacement_mapper([s, t], tok)


def test_word_inds_and_alpha_tables_match_reference(tok):
    from eta_inversion_b200.utils import ptp_utils
    k = 0
    for i, (s, t) in enumerate(H["pairs"]):
        assert [ptp_utils.get_word_inds(t, w, tok).tolist() for w in t.split(" ")] == H["word_inds"][i]
        for _ in range(2):
            g = H["alpha"][k]
            crs = g["crs"] if not isinstance(g["crs"], dict) else {kk: tuple(v) if isinstance(v, list) else v for kk, v in g["crs"].items()}
            al = ptp_utils.get_time_words_attention_alpha([s, t], 50, crs, tok)
            assert list(al.shape) == g["shape"] and al.reshape(51, -1).sum(1).tolist() == g["sum_per_step"]
            k += 1


def test_eta_schedule_matches_reference():
    from eta_inversion_b200.inversion.eta_inversion import make_eta_schedule
    for name, g in H["eta"].items():
        spec = g["spec"]
        spec = tuple(tuple(x) if isinstance(x, list) else x for x in spec) if isinstance(spec, list) else spec
        v = make_eta_schedule(spec)
        assert v.shape == (1000,)
        np.testing.assert_allclose(v[::37], np.array(g["values_every_37"]), rtol=1e-12, atol=1e-15)


def test_scheduler_constants_match_oracle():
    """Host half of the schedulers (alphas, timesteps, variance) vs the oracle restatement of diffusers' DDIMScheduler."""
    from eta_inversion_b200.models import sd_scheduler
    from eta_inversion_b200.inverse_schedulers import DDIMInverseScheduler, DDIMScheduler
    from oracle import sd15
    ours, ref = sd_scheduler(), sd15.sd_scheduler()
    assert torch.equal(ours.alphas_cumprod, ref.alphas_cumprod) and ours.config == ref.config
    for n in (50, 10, 4):
        a = DDIMScheduler.from_config({**ours.config, "clip_sample": False, "set_alpha_to_one": False})
        b = sd15.DDIMScheduler.from_config({**ref.config, "clip_sample": False, "set_alpha_to_one": False})
        a.set_timesteps(n), b.set_timesteps(n)
        assert torch.equal(a.timesteps, b.timesteps) and a.timesteps[-1] == 1
        inv = DDIMInverseScheduler.from_scheduler(a)
        inv.set_timesteps(n)
        assert inv.timesteps[0] == 1 and inv.timesteps[0] < inv.timesteps[1]
        for t in a.timesteps.tolist():
            p = a.prev_timestep(t)
            assert math.isclose(float(a._get_variance(t, p)), float(b._get_variance(t, p)), rel_tol=0, abs_tol=0)
            assert a.alpha(p) == float(b.alphas_cumprod[p] if p >= 0 else b.final_alpha_cumprod)


def test_psnr_known_answer_of_the_reference_metric_test():
    """test/test_metrics.py:61-62 of the reference: mse 0.011490068398416042, psnr 19.396774291992188 on its two PNGs."""
    import cv2
    from eta_inversion_b200.metrics import psnr
    from eta_inversion_b200.models import StablePreprocess
    pre = StablePreprocess("cpu", size=512)
    a = pre(GOLDEN / "ref_images" / "gnochi_mirror_sq.png")
    b = pre(GOLDEN / "ref_images" / "gnochi_mirror_sq_edit_example.png")
    assert a.shape == (1, 3, 512, 512) and -1 <= float(a.min()) and float(a.max()) <= 1
    assert abs(psnr(b, a) - 19.396774291992188) < 1e-4
    mse = torch.mean(((a + 1) / 2 - (b + 1) / 2) ** 2).item()
    assert abs(mse - 0.011490068398416042) < 1e-8


def test_flop_count_matches_survey():
    from eta_inversion_b200 import flops
    m = flops.unet_macs_per_row()
    assert abs(m["total"] / 1e9 - 401.637) < 0.01
    assert abs(m["conv3x3"] / m["total"] - 0.498) < 0.01


def test_param_spec_matches_oracle_module_tree():
    """The product's parameter spec (the loader contract for real diffusers checkpoints) names exactly the tensors of
    the oracle's UNet / VAE module trees."""
    from eta_inversion_b200 import synthetic as syn
    from oracle import sd15
    with torch.device("meta"):
        u, v = sd15.UNet2DConditionModel(), sd15.AutoencoderKL()
    assert {k: tuple(t.shape) for k, t in u.state_dict().items()} == dict(syn.unet_param_spec())
    assert {k: tuple(t.shape) for k, t in v.state_dict().items()} == dict(syn.vae_param_spec())
    assert sum(math.prod(s) for _, s in syn.unet_param_spec()) == 859_520_964


def test_sharding_covers_every_sample_once():
    from eta_inversion_b200.sweep import gather_results, group_indices, shard_indices
    for n, w in ((700, 8), (700, 4), (700, 2), (7, 8), (0, 2), (1, 1)):
        owned = [shard_indices(n, r, w) for r in range(w)]
        flat = sorted(i for o in owned for i in o)
        assert flat == list(range(n))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
    assert group_indices([0, 8, 16, 24, 32], 2) == [[0, 8], [16, 24], [32]]
    assert gather_results({0: "a", 1: "b"}, 2, 1) == ["a", "b"]
    with pytest.raises(RuntimeError):
        gather_results({0: "a"}, 2, 1)
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def test_ddpm_and_edict_inverse_schedulers_match_reference():
    """DDPMInverseScheduler (sampled trajectory, recovered noise maps, variance) and the EDICT scheduler pair (integer and
    fractional timestep spacing) against the reference's own classes (oracle/make_scheduler_goldens.py)."""
    import numpy as np
    from eta_inversion_b200.inverse_schedulers import DDIMScheduler, DDPMInverseScheduler
    from eta_inversion_b200.inversion.edict_inversion import EdictScheduler, EdictSchedulerInverse
    from eta_inversion_b200.models import sd_scheduler
    from pathlib import Path
    gold = np.load(Path(__file__).parent / "golden" / "schedulers.npz")
    z0, eps = torch.from_numpy(gold["z0"]), torch.from_numpy(gold["eps"])
    cfg = sd_scheduler().config
    for steps in (6, 3):
        for markov in (False, True):
            base = DDIMScheduler.from_config(cfg)
            base.set_timesteps(steps)
            inv = DDPMInverseScheduler.from_scheduler(base, markovian_forward=markov)
            inv.set_timesteps(steps)
            xts = inv.sample_latents(z0, generator=torch.Generator().manual_seed(5))
            key = f"ddpm_s{steps}_m{int(markov)}"
            assert np.abs(xts.numpy() - gold[key + "_xts"]).max() < 2e-6
            for i, t in enumerate(inv.timesteps):
                r = inv.step(eps[i % 6], t, inv.get_sampled_latent_by_t(xts, t), xts)
                scale = max(1.0, float(np.abs(gold[key + "_z"][i]).max()))
                assert np.abs(r.variance_noise.numpy() - gold[key + "_z"][i]).max() < 2e-5 * scale
                assert np.abs(r.prev_sample.numpy() - gold[key + "_x"][i]).max() < 2e-5
                assert abs(inv.get_variance(int(t)) - gold[key + "_var"][i]) < 1e-7
        base = DDIMScheduler.from_config(cfg)
        base.set_timesteps(steps)
        bwd, fwd = EdictScheduler(base), EdictSchedulerInverse(base)
        x = z0.clone()
        for i, t in enumerate(fwd.timesteps):
            x = fwd.step(eps[i % 6], t, x).prev_sample
            assert np.abs(x.numpy() - gold[f"edict_s{steps}_fwd"][i]).max() < 2e-5 * max(1.0, float(x.abs().max()))
        for i, t in enumerate(bwd.timesteps):
            x = bwd.step(eps[(steps - 1 - i) % 6], t, x).prev_sample
            assert np.abs(x.numpy() - gold[f"edict_s{steps}_bwd"][i]).max() < 2e-5 * max(1.0, float(x.abs().max()))


def test_attn_control_drop_leading_rows():
    """AttnControl.drop_leading_rows: the control of a CFG batch whose unconditional half is not computed (EtaInversion at
    guidance_scale_fwd == 1) -- rows renumbered, None when the control reads or writes a dropped row."""
    from eta_inversion_b200.engine import AttnControl
    store = AttnControl(store_rows=[2, 3], store_res=16)
    half = store.drop_leading_rows(2)
    assert half.store_rows == [0, 1] and half.store_res == 16 and half.self_rows is None and half.edit_pairs is None
    assert AttnControl(store_rows=[1, 3]).drop_leading_rows(2) is None            # stores an unconditional row
    remap = AttnControl(self_rows=([0, 1, 2, 2], [0, 1, 2, 2], [0, 1, 2, 3]), self_layer_mask=0xFF, self_max_tokens=256)
    h = remap.drop_leading_rows(2)
    assert h.self_rows == ([0, 0], [0, 0], [0, 1]) and h.self_layer_mask == 0xFF and h.self_max_tokens == 256
    masa = AttnControl(self_rows=([0, 1, 2, 3], [0, 0, 2, 2], [0, 0, 2, 2]))      # MasaCtrl: rows of each half read the half's source
    assert masa.drop_leading_rows(2).self_rows == ([0, 1], [0, 0], [0, 0])
    assert AttnControl(self_rows=([0, 1, 0, 1], [0, 1, 2, 3], [0, 1, 2, 3])).drop_leading_rows(2) is None  # reads a dropped row
    assert AttnControl(edit_pairs=[(2, 3)]).drop_leading_rows(2).edit_pairs == [(0, 1)]
    assert AttnControl(edit_pairs=[(0, 3)]).drop_leading_rows(2) is None
    assert AttnControl(conv_inject_rows=1).drop_leading_rows(1) is None


def test_null_text_loss_gradient_closed_form():
    """The closed-form dL/d(eps_uncond) NullTextInversion feeds to etai_unet_backward_ctx equals torch autograd through the
    reference's own arithmetic: CFG combine, DDIMScheduler.step (eta = 0), mse_loss (null_text_inversion.py:72-76)."""
    from oracle import sd15
    sch = sd15.sd_scheduler()
    sch.set_timesteps(10)
    g = torch.Generator().manual_seed(3)
    x, e_c, target = (torch.randn((1, 4, 64, 64), generator=g) for _ in range(3))
    e_u = torch.randn((1, 4, 64, 64), generator=g).requires_grad_(True)
    gs = 7.5
    for t in (sch.timesteps[0], sch.timesteps[4], sch.timesteps[-1]):
        e_u.grad = None
        eps = e_u + gs * (e_c - e_u)
        rec = sch.step(eps, t, x).prev_sample
        loss = torch.nn.functional.mse_loss(rec, target)
        loss.backward()
        a_t = float(sch.alphas_cumprod[int(t)])
        p = int(t) - sch.config.num_train_timesteps // sch.num_inference_steps
        a_p = float(sch.alphas_cumprod[p]) if p >= 0 else float(sch.final_alpha_cumprod)
        coef = (1.0 - gs) * ((1.0 - a_p) ** 0.5 - (a_p ** 0.5) * ((1.0 - a_t) ** 0.5) / (a_t ** 0.5))
        closed = (rec.detach() - target) * (2.0 * coef / rec.numel())
        assert torch.allclose(e_u.grad, closed, rtol=2e-4, atol=1e-12), float((e_u.grad - closed).abs().max())


def test_tensor_table_converts_plain_tensors_on_the_host():
    """Weight table handed to etai_*_create: vectors, matrices and 1x1 convolutions are converted to the storage dtype on
    the host (the library then only copies them); 3x3 filters and the GEGLU projection stay fp32 (re-packed on the device)."""
    from eta_inversion_b200 import _lib
    sd = {"a.norm.weight": torch.ones(8), "a.to_q.weight": torch.randn(8, 8), "a.proj_in.weight": torch.randn(8, 8, 1, 1),
          "a.conv1.weight": torch.randn(8, 8, 3, 3), "b.ff.net.0.proj.weight": torch.randn(64, 8), "b.ff.net.0.proj.bias": torch.randn(64),
          "ids": torch.arange(4)}
    arr, keep = _lib.tensor_table(sd, torch.float16)
    got = {arr[i].name.decode(): (arr[i].dtype, keep[i]) for i in range(len(sd))}
    for name in ("a.norm.weight", "a.to_q.weight", "a.proj_in.weight"):
        assert got[name][0] == _lib.ETAI_F16 and got[name][1].dtype == torch.float16
        assert torch.equal(got[name][1], sd[name].to(torch.float16))
    for name in ("a.conv1.weight", "b.ff.net.0.proj.weight", "b.ff.net.0.proj.bias"):
        assert got[name][0] == _lib.ETAI_F32 and got[name][1].dtype == torch.float32
    assert got["ids"][1].dtype == torch.float16            # integer tensors are widened to float first, then plain
    arr32, keep32 = _lib.tensor_table(sd)                    # no storage dtype: nothing is converted
    assert all(k.dtype == torch.float32 for k in keep32)


def test_quantile_ranks_reproduce_torch_quantile():
    """engine.quantile_ranks feeds etai_prox_guidance: lerp(sorted[lo], sorted[hi], w) must be torch.quantile's 'linear'
    result bit for bit (float32 rank arithmetic of ATen's quantile_compute), including integer ranks and both ends."""
    import torch
    from eta_inversion_b200.engine import quantile_ranks
    g = torch.Generator().manual_seed(0)
    for n in (7, 100, 1000, 16384, 32768, 49152):
        x = torch.randn(n, generator=g).abs()
        s = x.sort().values
        for q in (0.0, 0.0001, 0.25, 0.3, 0.5, 0.7, 0.999, 1.0):
            lo, hi, w = quantile_ranks(q, n)
            assert 0 <= lo <= hi <= min(lo + 1, n - 1) and 0.0 <= w < 1.0
            assert torch.lerp(s[lo], s[hi], torch.tensor(w)) == x.quantile(q), (n, q)
