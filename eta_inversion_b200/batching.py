"""Lock-step co-batching of independent edits on one GPU.

The reference runs one edit at a time per process (eval.py:65-109), which at B200 speed leaves the UNet forward
latency-bound: at B=2 / B=4 most of the ~420 kernels of a forward are far too small to fill 148 SMs.  Edits are
independent, so k of them can walk their loops in lock step and share every UNet forward:
each edit ("lane") keeps its own unmodified inverter / editor / controller objects and runs in its own Python thread;
the lanes' ``unet(...)`` calls rendez-vous here, are concatenated along the batch dimension (rows of lane l are
[l*B, (l+1)*B)), their attention-control descriptors are merged (row indices shifted, PtP tables stacked, one pair per
lane) and ONE engine forward serves all lanes.  Nothing else is shared: schedulers, noise, LocalBlend and the VAE
stay per lane, so results equal the sequential ones (bit-exact on the fp32 path, whose kernels are batch-invariant).
"""
from __future__ import annotations

import copy
import sys
import threading
from typing import Any, Callable, Dict, List, Optional, Sequence

import torch

from .engine import AttnControl, UNetEngine, UNetOutput


class _FastGilSwitch:
    """Lanes hand the GIL over at every rendez-vous; the interpreter's 5 ms default switch interval stalls the others.
    The setting is process-wide and groups run concurrently (run_pipelined), so it is reference-counted: the first group in
    shortens it, the last one out restores what it found."""

    _lock, _users, _saved = threading.Lock(), 0, None

    def __enter__(self):
        cls = _FastGilSwitch
        with cls._lock:
            if cls._users == 0:
                cls._saved = sys.getswitchinterval()
                sys.setswitchinterval(2e-4)
            cls._users += 1

    def __exit__(self, *exc):
        cls = _FastGilSwitch
        with cls._lock:
            cls._users -= 1
            if cls._users == 0:
                sys.setswitchinterval(cls._saved)


class _LaneUNet:
    """What lane `lane` sees as ``pipe.unet``."""

    def __init__(self, group: "LockstepGroup", lane: int) -> None:
        self._group, self._lane = group, lane

    def __getattr__(self, name):  # dtype, device, latent_hw, launch_count, ...
        return getattr(self._group.engine, name)

    def __call__(self, sample, timestep, encoder_hidden_states=None, control=None, **kwargs) -> UNetOutput:
        return self._group.forward(self._lane, sample, timestep, encoder_hidden_states, control)


class LockstepGroup:
    """Rendez-vous of the lanes' ``unet(...)`` calls.  The set of LIVE lanes shrinks when a lane finishes (an edit that
    returns None because it is unsupported -- eta_inversion.py:385-386 of the reference --, or one that simply takes fewer
    UNet forwards than its peers): the remaining lanes keep walking in lock step with a smaller merged batch instead of
    waiting for a request that will never come."""

    def __init__(self, engine: UNetEngine, lanes: int, timeout_s: float = 600.0) -> None:
        self.engine, self.lanes, self.timeout = engine, lanes, timeout_s
        self._cv = threading.Condition()
        self._live = set(range(lanes))
        self._req: Dict[int, Any] = {}
        self._out: Dict[int, Any] = {}
        self._gen = 0
        self._err: Optional[BaseException] = None
        self._aborted = False
        self._ctx_key, self._ctx_cat, self._ctx_role_major = None, None, False
        self._table_cache: Dict[str, Any] = {}
        self._store_cache = None
        self.forwards = 0

    def lane_unet(self, lane: int) -> _LaneUNet:
        return _LaneUNet(self, lane)

    def abort(self) -> None:
        with self._cv:
            self._aborted = True
            self._cv.notify_all()

    def lane_done(self, lane: int) -> None:
        """Lane `lane` will post no more requests; if the others were only waiting for it, their forward runs now."""
        with self._cv:
            self._live.discard(lane)
            if self._live and not self._aborted and len(self._req) == len(self._live):
                self._run_locked()

    # ---- called concurrently by the lane threads ---------------------------------------------------
    def forward(self, lane: int, sample, timestep, ctx, control) -> UNetOutput:
        with self._cv:
            if self._aborted:
                raise RuntimeError("etai lockstep: group aborted (another lane failed)") from self._err
            self._req[lane] = (sample, timestep, ctx, control)
            gen = self._gen
            if len(self._req) == len(self._live):  # every live lane has posted its request: this thread runs the engine
                self._run_locked()
            elif not self._cv.wait_for(lambda: self._gen != gen or self._aborted, self.timeout):
                self._aborted = True
                self._cv.notify_all()
                raise RuntimeError(f"etai lockstep: lane {lane} waited {self.timeout:.0f} s for its peers")
            if self._err is not None:
                raise RuntimeError(f"etai lockstep: batched forward failed: {self._err}") from self._err
            if self._aborted:
                raise RuntimeError("etai lockstep: group aborted (another lane failed)")
            return UNetOutput(self._out.pop(lane))

    # ---- leader (holds the condition's lock; the other lanes are parked in wait_for) ----------------
    def _run_locked(self) -> None:
        try:
            self._run()
        except BaseException as e:  # noqa: BLE001 - re-raised in every lane
            self._err = e
            self._aborted = True
        self._req = {}
        self._gen += 1
        self._cv.notify_all()

    def _run(self) -> None:
        order = sorted(self._req)          # live lanes, in lane order: rows of lane order[i] are [i*B, (i+1)*B)
        reqs = [self._req[l] for l in order]
        B = reqs[0][0].shape[0]
        t0 = float(reqs[0][1])
        for s, t, c, _ in reqs:
            if s.shape[0] != B or float(t) != t0 or c is None or c.shape[0] != B:
                raise RuntimeError("lanes disagree on batch rows / timestep / context rows")
        L = len(reqs)
        # Row layout of the merged batch.  Lane-major (lane l owns rows [l*B, (l+1)*B)) unless a lane asks for Plug-and-Play's
        # feature injection: the engine copies rows [0, n) over rows [n, 3n) (pnp_utils.py:172-177), so the lanes' source
        # rows must come first -- role-major, row r of lane l at r*L + l.
        role_major = any(r[3] is not None and r[3].conv_inject_rows for r in reqs)
        if role_major:
            sample = torch.stack([r[0] for r in reqs], 1).flatten(0, 1)
        else:
            sample = torch.cat([r[0] for r in reqs])
        old = self._ctx_key  # strong references to the lanes' context tensors + their versions (addresses can be reused)
        same = (old is not None and len(old) == L and self._ctx_role_major == role_major
                and all(o[0] is r[2] and o[1] == r[2]._version for o, r in zip(old, reqs)))
        if not same:  # contexts are constant over a loop: concatenate (and re-project K/V) only on change
            self._ctx_key = [(r[2], r[2]._version) for r in reqs]
            self._ctx_role_major = role_major
            self._ctx_cat = (torch.stack([r[2] for r in reqs], 1).flatten(0, 1) if role_major
                             else torch.cat([r[2] for r in reqs])).contiguous()
        ctrl, scatter = self._merge([r[3] for r in reqs], B, role_major)
        eps = self.engine(sample, t0, encoder_hidden_states=self._ctx_cat, control=ctrl)["sample"]
        for fn in scatter:
            fn()
        for i, l in enumerate(order):
            self._out[l] = eps[i::L] if role_major else eps[i * B:(i + 1) * B]
        self.forwards += 1

    def _merge(self, ctrls: Sequence[Optional[AttnControl]], B: int, role_major: bool = False):
        if all(c is None for c in ctrls):
            return None, []
        out = AttnControl()
        scatter: List[Callable[[], None]] = []
        L = len(ctrls)
        rm = (lambda l, r: r * L + l) if role_major else (lambda l, r: l * B + r)  # merged row of row r of lane l
        # self-attention remap: lanes without one keep identity rows
        if any(c is not None and c.self_rows is not None for c in ctrls):
            masks = {(c.self_layer_mask, c.self_max_tokens) for c in ctrls if c is not None and c.self_rows is not None}
            if len(masks) != 1:
                raise RuntimeError("lanes disagree on the self-attention remap window")
            out.self_layer_mask, out.self_max_tokens = masks.pop()
            q, k, v = [0] * (L * B), [0] * (L * B), [0] * (L * B)
            for l, c in enumerate(ctrls):
                rows = c.self_rows if (c is not None and c.self_rows is not None) else ([*range(B)],) * 3
                for r in range(B):
                    q[rm(l, r)], k[rm(l, r)], v[rm(l, r)] = rm(l, rows[0][r]), rm(l, rows[1][r]), rm(l, rows[2][r])
            out.self_rows = (q, k, v)
        # cross-attention edit: one (base, target) pair per lane that edits at this step
        ed = [(l, c) for l, c in enumerate(ctrls) if c is not None and c.edit_pairs is not None]
        if ed:
            out.edit_pairs = [(rm(l, b), rm(l, t)) for l, c in ed for (b, t) in c.edit_pairs]
            for name in ("mapper", "blend_a", "equalizer", "alpha_step"):
                parts = [getattr(c, name) for _, c in ed]
                # mapper / blend_a / equalizer are the same tensor objects for a whole loop: concatenate once, not per forward
                # (the lanes' Python runs under one GIL, every op saved here is saved on the group's critical path)
                key = (name, tuple((id(t), t._version) for t in parts))
                hit = self._table_cache.get(name)
                if hit is None or hit[0] != key:
                    hit = (key, torch.cat(parts).contiguous(), parts)  # parts: strong references keep the ids valid
                    self._table_cache[name] = hit
                setattr(out, name, hit[1])
        # attention store: one combined accumulator per place, scattered back to the lanes' own tensors afterwards
        st = [(l, c) for l, c in enumerate(ctrls) if c is not None and c.store_rows is not None]
        if st:
            res = {c.store_res for _, c in st}
            if len(res) != 1:
                raise RuntimeError("lanes disagree on the attention-store resolution")
            out.store_res = res.pop()
            out.store_rows = [rm(l, r) for l, c in st for r in c.store_rows]
            n = len(out.store_rows)
            places = [p for p in ("store_down", "store_mid", "store_up") if any(getattr(c, p) is not None for _, c in st)]
            # one combined accumulator per place, kept across forwards and re-zeroed with ONE multi-tensor launch; the lanes'
            # own accumulators receive their slices with ONE multi-tensor add after the forward (was: 3 allocations + memsets
            # and one add per lane and place, 15 launches per forward for 4 lanes)
            ckey = (n, out.store_res, tuple(places))
            if self._store_cache is None or self._store_cache[0] != ckey:
                self._store_cache = (ckey, [torch.zeros((n, out.store_res ** 2, 77), dtype=torch.float32,
                                                        device=self.engine.device) for _ in places])
            else:
                torch._foreach_zero_(self._store_cache[1])
            dsts, srcs = [], []
            for place, acc in zip(places, self._store_cache[1]):
                setattr(out, place, acc)
                off = 0
                for _, c in st:
                    m = len(c.store_rows)
                    dst = getattr(c, place)
                    if dst is not None:
                        dsts.append(dst)
                        srcs.append(acc[off:off + m])
                    off += m
            if dsts:
                scatter.append(lambda dsts=dsts, srcs=srcs: torch._foreach_add_(dsts, srcs))
        if role_major:
            # every lane injects its row 0 into its rows 1, 2: with the source rows first this is ONE block copy in the engine
            if any(c is None or c.conv_inject_rows != 1 for c in ctrls) or B != 3:
                raise RuntimeError("lanes disagree on the Plug-and-Play feature injection (every lane must inject 1 source row "
                                   "into its 3-row batch at this step)")
            out.conv_inject_rows = L
        return out, scatter


def run_lockstep(pipe, jobs: Sequence[Dict[str, Any]], make_editor: Callable[[Any], Any],
                 stream: Optional[torch.cuda.Stream] = None) -> List[Any]:
    """Run ``len(jobs)`` edits in lock step on ``pipe``.  ``stream``: CUDA stream every lane issues its work on
    (default: each thread's current stream); groups that run concurrently must use different streams.

    jobs[i]: kwargs of ``Editor.edit`` (image, source_prompt, target_prompt, cfg, inv_cfg).
    make_editor(lane_pipe) -> editor: builds the lane's own inverter + editor on a per-lane view of the pipeline
    (same VAE / text encoder / tokenizer, its own scheduler slot and the lock-step UNet proxy).
    The engine behind ``pipe.unet`` must have been created with ``max_batch >= 4 * len(jobs)``.
    """
    k = len(jobs)
    group = LockstepGroup(pipe.unet, k)
    editors = []
    with torch.cuda.stream(stream):  # constructors may create device tensors: same stream as the lanes' work
        for l in range(k):
            lane_pipe = copy.copy(pipe)
            lane_pipe.unet = group.lane_unet(l)
            editors.append(make_editor(lane_pipe))
    results: List[Any] = [None] * k
    errors: List[Optional[BaseException]] = [None] * k
    dev = pipe.device

    def work(l: int) -> None:
        try:
            torch.cuda.set_device(dev)
            with torch.no_grad(), torch.cuda.stream(stream):  # stream=None is a no-op context
                results[l] = editors[l].edit(**jobs[l])
            group.lane_done(l)  # the peers go on without this lane (it may have returned early, e.g. None = unsupported)
        except BaseException as e:  # noqa: BLE001
            errors[l] = e
            group.abort()

    threads = [threading.Thread(target=work, args=(l,), name=f"etai-lane-{l}") for l in range(k)]
    with _FastGilSwitch():
        for th in threads:
            th.start()
        for th in threads:
            th.join()
    for e in errors:  # the root cause first: lanes that were merely aborted report "group aborted"
        if e is not None and "group aborted" not in str(e):
            raise e
    for e in errors:
        if e is not None:
            raise e
    return results


def run_pipelined(pipes: Sequence[Any], job_groups: Sequence[Sequence[Dict[str, Any]]],
                  make_editor: Callable[[Any], Any],
                  on_result: Optional[Callable[[int, List[Any]], Any]] = None) -> List[List[Any]]:
    """Run several lock-step groups with ``len(pipes)`` of them in flight at any time.

    One group alone leaves the GPU idle while its lanes set up (text encoder, VAE encode, controllers: ~160 ms per
    group of 4), while the host drains the launch queue between the inversion and the editing loop, and during the
    final VAE decode / device-to-host copy -- about 10 % of a 50-step edit at B200 speed.  Edits are independent
    (SURVEY.md section 8e), so a second group on its OWN engine (own activation arena, CUDA graphs and stream; weights
    are read-only) fills those gaps: group g runs on ``pipes[g % len(pipes)]``, each pipe is driven by one thread.
    Results are returned in ``job_groups`` order.  Tensors in the jobs must be complete before the call (the groups
    run on side streams that only synchronise with the caller at entry and exit); images may be pinned host tensors
    (each lane uploads its own).  ``on_result(g, results)`` runs in the driving thread as soon as group g is complete
    (e.g. device-to-host copies / file writes, which then overlap the other pipe's work); its return value replaces
    the group's results."""
    n = len(pipes)
    out: List[Any] = [None] * len(job_groups)
    errors: List[Optional[BaseException]] = [None] * n
    dev = pipes[0].device
    torch.cuda.synchronize(dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(n)]

    def drive(w: int) -> None:
        try:
            torch.cuda.set_device(dev)
            for g in range(w, len(job_groups), n):
                out[g] = run_lockstep(pipes[w], job_groups[g], make_editor, stream=streams[w])
                streams[w].synchronize()  # results of group g are complete; its buffers may be reused by group g + n
                if on_result is not None:
                    out[g] = on_result(g, out[g])
        except BaseException as e:  # noqa: BLE001
            errors[w] = e

    threads = [threading.Thread(target=drive, args=(w,), name=f"etai-pipe-{w}") for w in range(n)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for e in errors:
        if e is not None:
            raise e
    return out
