"""Host logic of the lock-step co-batching (no GPU): how the lanes' attention-control descriptors are merged into one
descriptor for the shared forward (row indices shifted by lane, PtP tables stacked, one store accumulator scattered
back to the lanes), and the rendez-vous itself with a stub engine."""
import threading
from types import SimpleNamespace

import pytest
import torch

from eta_inversion_b200.batching import LockstepGroup
from eta_inversion_b200.engine import AttnControl


class _StubEngine:
    """Records what the group hands to the engine; eps = sample + row index so lanes can check their slice."""
    device = torch.device("cpu")

    def __init__(self):
        self.calls = []

    def __call__(self, sample, t, encoder_hidden_states=None, control=None):
        self.calls.append((sample.clone(), float(t), encoder_hidden_states.clone(), control))
        if control is not None and control.store_rows is not None:
            for place in ("store_down", "store_mid", "store_up"):
                acc = getattr(control, place)
                if acc is not None:
                    acc += torch.arange(1, acc.shape[0] + 1, dtype=torch.float32).reshape(-1, 1, 1)  # row r gets r + 1
        return {"sample": sample + torch.arange(sample.shape[0], dtype=sample.dtype).reshape(-1, 1, 1, 1)}


def _ptp_ctrl(val):
    c = AttnControl()
    c.self_rows = ([0, 1, 2, 2], [0, 1, 2, 2], [0, 1, 2, 3])
    c.self_layer_mask, c.self_max_tokens = 0xFFFF, 1024
    c.edit_pairs = [(2, 3)]
    c.mapper = torch.full((1, 77, 77), float(val))
    c.blend_a, c.equalizer, c.alpha_step = torch.full((1, 77), val + .1), torch.full((1, 77), val + .2), torch.full((1, 77), val + .3)
    c.store_rows, c.store_res = [2, 3], 16
    c.store_down, c.store_up = torch.zeros((2, 256, 77)), torch.zeros((2, 256, 77))
    return c


def test_merge_shifts_rows_and_stacks_tables():
    eng = _StubEngine()
    grp = LockstepGroup(eng, lanes=3)
    ctrls = [_ptp_ctrl(1.0), None, _ptp_ctrl(3.0)]  # the middle lane runs plain attention at this step
    merged, scatter = grp._merge(ctrls, B=4)
    assert merged.self_rows[0] == [0, 1, 2, 2, 4, 5, 6, 7, 8, 9, 10, 10]       # lane 1 keeps identity rows
    assert merged.self_rows[2] == [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11]
    assert merged.edit_pairs == [(2, 3), (10, 11)]
    assert merged.mapper.shape == (2, 77, 77) and merged.mapper[0, 0, 0] == 1.0 and merged.mapper[1, 0, 0] == 3.0
    assert torch.allclose(merged.alpha_step[:, 0], torch.tensor([1.3, 3.3]))
    assert merged.store_rows == [2, 3, 10, 11] and merged.store_down.shape == (4, 256, 77) and merged.store_mid is None
    merged.store_down += torch.arange(1, 5, dtype=torch.float32).reshape(-1, 1, 1)
    for fn in scatter:
        fn()
    assert ctrls[0].store_down[:, 0, 0].tolist() == [1.0, 2.0] and ctrls[2].store_down[:, 0, 0].tolist() == [3.0, 4.0]
    assert grp._merge([None, None, None], 4) == (None, [])


def test_merge_rejects_inconsistent_lanes():
    grp = LockstepGroup(_StubEngine(), lanes=2)
    a, b = _ptp_ctrl(1.0), _ptp_ctrl(2.0)
    b.self_max_tokens = 256
    with pytest.raises(RuntimeError, match="self-attention remap window"):
        grp._merge([a, b], 4)
    b = _ptp_ctrl(2.0)
    b.store_res = 32
    with pytest.raises(RuntimeError, match="attention-store resolution"):
        grp._merge([a, b], 4)
    b = _ptp_ctrl(2.0)
    b.conv_inject_rows = 1  # one lane injects Plug-and-Play features, the other does not (and the batch is not PnP's 3 rows)
    with pytest.raises(RuntimeError, match="Plug-and-Play"):
        grp._merge([a, b], 4, role_major=True)


def test_lanes_rendezvous_into_one_forward_and_get_their_rows_back():
    eng = _StubEngine()
    lanes = 3
    grp = LockstepGroup(eng, lanes, timeout_s=30)
    out = [None] * lanes

    def work(l):
        unet = grp.lane_unet(l)
        x = torch.full((2, 4, 8, 8), float(10 * l))
        ctx = torch.full((2, 77, 8), float(l))
        for step in range(3):
            x = unet(x, 981 - step, encoder_hidden_states=ctx, control=None)["sample"]
        out[l] = x
    ths = [threading.Thread(target=work, args=(l,)) for l in range(lanes)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert len(eng.calls) == 3 and grp.forwards == 3                      # one engine forward per step for all lanes
    assert eng.calls[0][0].shape[0] == 6 and eng.calls[0][2][:, 0, 0].tolist() == [0, 0, 1, 1, 2, 2]
    for l in range(lanes):  # each step added the lane's global row index (2l, 2l+1)
        assert out[l][:, 0, 0, 0].tolist() == [10 * l + 3 * (2 * l), 10 * l + 3 * (2 * l + 1)]


def test_lanes_that_disagree_fail_loudly():
    eng = _StubEngine()
    grp = LockstepGroup(eng, 2, timeout_s=10)
    errs = []

    def work(l):
        try:
            grp.lane_unet(l)(torch.zeros((2, 4, 8, 8)), 981 - l, encoder_hidden_states=torch.zeros((2, 77, 8)))  # different t
        except RuntimeError as e:
            errs.append(str(e))
    ths = [threading.Thread(target=work, args=(l,)) for l in range(2)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert len(errs) == 2 and all("lanes disagree" in e for e in errs)


def test_a_lane_that_finishes_early_releases_its_peers():
    """An edit that returns early (EtaInversion.invert -> None for an unsupported sample, eta_inversion.py:385-386, or
    a loop with fewer UNet forwards) must not strand the other lanes: they go on with a smaller merged batch."""
    from eta_inversion_b200 import batching

    eng = _StubEngine()
    pipe = SimpleNamespace(unet=eng, device=torch.device("cpu"))

    class _Editor:
        def __init__(self, p):
            self.unet = p.unet

        def edit(self, lane, steps):
            if steps == 0:
                return None                     # "unsupported": no UNet forward at all
            x = torch.full((2, 4, 8, 8), float(lane))
            ctx = torch.full((2, 77, 8), float(lane))
            for s in range(steps):
                x = self.unet(x, 981 - s, encoder_hidden_states=ctx)["sample"]
            return x

    orig_set_device = torch.cuda.set_device
    torch.cuda.set_device = lambda d: None      # CPU-only test: run_lockstep pins the device in every lane thread
    try:
        res = batching.run_lockstep(pipe, [dict(lane=0, steps=3), dict(lane=1, steps=0), dict(lane=2, steps=2)], _Editor)
    finally:
        torch.cuda.set_device = orig_set_device
    assert res[1] is None
    # steps 0,1: lanes 0 and 2 share a B=4 forward; step 2: lane 0 alone (B=2)
    assert [c[0].shape[0] for c in eng.calls] == [4, 4, 2]
    assert res[0][:, 0, 0, 0].tolist() == [0.0, 3.0]            # rows 0,1 in every forward
    assert res[2][:, 0, 0, 0].tolist() == [2.0 + 2 * 2, 2.0 + 2 * 3]  # rows 2,3 of the two shared forwards


def test_merge_role_major_for_plug_and_play():
    """Plug-and-Play lanes (3 rows: source, uncond, cond; row 0's features and q/k are injected into rows 1, 2) are merged
    role-major -- all sources first -- so that the engine's block copy rows [0, n) -> [n, 3n) serves every lane."""
    eng = _StubEngine()
    grp = LockstepGroup(eng, lanes=2)

    def pnp_ctrl():
        c = AttnControl(conv_inject_rows=1)
        c.self_rows = ([0, 0, 0], [0, 0, 0], [0, 1, 2])
        c.self_layer_mask = 0xFF00
        return c
    merged, _ = grp._merge([pnp_ctrl(), pnp_ctrl()], B=3, role_major=True)
    assert merged.conv_inject_rows == 2
    # merged row of (lane l, row r) = r * 2 + l: rows [src0, src1, unc0, unc1, cond0, cond1]
    assert merged.self_rows[0] == [0, 1, 0, 1, 0, 1] and merged.self_rows[1] == [0, 1, 0, 1, 0, 1]
    assert merged.self_rows[2] == [0, 1, 2, 3, 4, 5] and merged.self_layer_mask == 0xFF00
    with pytest.raises(RuntimeError, match="Plug-and-Play"):
        grp._merge([pnp_ctrl(), AttnControl()], B=3, role_major=True)
    # the rendez-vous lays samples / contexts out the same way and hands every lane its own rows back
    outs = [None, None]

    def lane(l):
        x = torch.full((3, 4, 8, 8), float(10 * l)) + torch.arange(3, dtype=torch.float32).reshape(3, 1, 1, 1)
        outs[l] = grp.forward(l, x, 500, torch.full((3, 77, 768), float(l)), pnp_ctrl())["sample"]
    ths = [threading.Thread(target=lane, args=(l,)) for l in range(2)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    sample, _, ctx, ctrl = eng.calls[-1]
    assert sample[:, 0, 0, 0].tolist() == [0.0, 10.0, 1.0, 11.0, 2.0, 12.0] and ctx[:, 0, 0].tolist() == [0, 1, 0, 1, 0, 1]
    assert ctrl.conv_inject_rows == 2
    # stub: eps = sample + merged row index -> lane 0 rows 0, 2, 4 ; lane 1 rows 1, 3, 5
    assert outs[0][:, 0, 0, 0].tolist() == [0.0, 3.0, 6.0] and outs[1][:, 0, 0, 0].tolist() == [11.0, 14.0, 17.0]


def test_gil_switch_interval_is_restored_after_overlapping_groups():
    """run_lockstep shortens the interpreter's switch interval while lanes run; groups overlap in run_pipelined, so the
    process-wide setting is reference-counted: inner exits must not restore early, the last exit restores the original."""
    import sys
    from eta_inversion_b200.batching import _FastGilSwitch
    before = sys.getswitchinterval()
    a, b = _FastGilSwitch(), _FastGilSwitch()
    a.__enter__()
    fast = sys.getswitchinterval()
    assert fast < before
    b.__enter__()
    a.__exit__(None, None, None)
    assert sys.getswitchinterval() == fast       # group b is still running
    b.__exit__(None, None, None)
    assert sys.getswitchinterval() == before
