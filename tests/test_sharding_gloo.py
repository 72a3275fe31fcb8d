"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: item sharding by trajectory, scatter, frame gather, max-time
reduction.  The denoising itself is replaced by a deterministic stand-in: the N-rank result set must equal the 1-rank set."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mudg_b200 import shard   # noqa: E402


def fake_clip(item):
    g = torch.Generator().manual_seed(123 + item["id"])
    return torch.rand(4, 3, 8, 8, generator=g) * 2 - 1


def make_items():
    # 3 trajectories x 2 windows + 2 single-window trajectories
    items, i = [], 0
    for traj in ("a", "b", "c"):
        for wdw in range(2):
            items.append({"id": i, "traj": traj, "window": wdw}); i += 1
    for traj in ("d", "e"):
        items.append({"id": i, "traj": traj, "window": 0}); i += 1
    return items


def _worker(rank, ws, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        mine = shard.scatter_items(make_items() if rank == 0 else None, key=lambda it: it["traj"])
        # a trajectory never straddles ranks and its windows stay ordered
        trajs = {it["traj"] for it in mine}
        frames = torch.stack([shard.frames_to_uint8(fake_clip(it)) for it in mine]) if mine else torch.zeros((0, 4, 3, 8, 8), dtype=torch.uint8)
        ids = torch.tensor([it["id"] for it in mine], dtype=torch.long)
        allf, alli = shard.gather_frames(frames, ids)
        tmax = shard.max_over_ranks(1.0 + rank)
        # gather to the writing rank only (bench.py's e2e leg): equal clip counts per call
        one = shard.frames_to_uint8(fake_clip({"id": 100 + rank}))[None]
        gf, gi = shard.gather_frames_to(one, torch.tensor([100 + rank]), dst=0)
        if rank == 0:
            assert gi.tolist() == [100, 101] and torch.equal(gf[1], shard.frames_to_uint8(fake_clip({"id": 101})))
        else:
            assert gf is None and gi is None
        q.put((rank, sorted(trajs), [(it["traj"], it["window"]) for it in mine], allf.clone(), alli.clone(), tmax))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_equals_single_rank():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda r: r[0])
    t0, t1 = set(res[0][1]), set(res[1][1])
    assert not (t0 & t1) and (t0 | t1) == {"a", "b", "c", "d", "e"}
    for _, _, order, _, _, _ in res:
        for traj in {t for t, _ in order}:
            ws_ = [w for t, w in order if t == traj]
            assert ws_ == sorted(ws_)
    # single-rank reference
    items = make_items()
    ref = torch.stack([shard.frames_to_uint8(fake_clip(it)) for it in items])
    for _, _, _, allf, alli, tmax in res:
        assert alli.tolist() == list(range(len(items)))
        assert torch.equal(allf, ref)                      # bitwise: clips are independent
        assert tmax == 2.0


def test_assign_items_round_robin():
    items = list(range(10))
    assert shard.assign_items(items, 0, 4) == [0, 4, 8]
    assert shard.assign_items(items, 3, 4) == [3, 7]
    assert sum(len(shard.assign_items(items, r, 4)) for r in range(4)) == 10


def test_frames_to_uint8_is_the_reference_conversion():
    """Truncation, not rounding: bit for bit the conversion pinned in oracle/post_oracle.py against eval_tools.py:24-27."""
    import numpy as np
    from oracle import post_oracle as P
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 4, 8, 8, generator=g)
    x[0, 0, 0, 0, :4] = torch.tensor([-1.0, 1.0, 1.0 / 255 * 2 - 1, 0.5])
    assert np.array_equal(shard.frames_to_uint8(x).numpy(), P.to_uint8(x.numpy()))
