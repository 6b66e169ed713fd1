"""World-size-2 test (gloo, CPU) of the multi-GPU path's host logic: per-image sharding, lock-step grouping and the
end-of-sweep gather.  The GPU run uses the same code with the nccl backend (bench.py / eval.py)."""
import os

import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from eta_inversion_b200.sweep import run_sweep
    calls = []

    def run_group(idx):
        calls.append(list(idx))
        return [{"sample": i, "rank": rank, "checksum": (i * 2654435761) % 997} for i in idx]
    res = run_sweep(23, rank, world, cobatch=4, run_group=run_group)
    assert [r["sample"] for r in res] == list(range(23))
    assert all(r["rank"] == r["sample"] % world for r in res)
    assert all(len(c) <= 4 for c in calls) and sum(len(c) for c in calls) == len(range(rank, 23, world))
    # windowed form (several lock-step groups handed over at once, as eval.py does for batching.run_pipelined)
    windows = []

    def run_groups(groups):
        windows.append([list(g) for g in groups])
        return [run_group(g) for g in groups]
    res2 = run_sweep(23, rank, world, cobatch=4, run_group=run_group, window=2, run_groups=run_groups)
    assert res2 == res
    assert all(len(w) <= 2 for w in windows) and sum(len(g) for w in windows for g in w) == len(range(rank, 23, world))
    out[rank] = [r["checksum"] for r in res]
    dist.destroy_process_group()


def test_two_rank_sweep_gathers_identical_ordered_results():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        assert out[0] == out[1] and len(out[0]) == 23
