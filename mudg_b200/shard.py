"""Multi-GPU plumbing of the sampler: independent clips are sharded over ranks, there is NO collective on the denoising
path.  torch.distributed (NCCL on GPUs, gloo in CPU tests) only scatters the item list, gathers the
decoded uint8 frames and reduces timings."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def assign_items(items: Sequence, rank: int, world_size: int, key=None) -> List:
    """Round-robin by trajectory: consecutive windows of one trajectory depend on each other (8-frame overlap feedback,
    virtual_pose_render.py:262-274) so all items with the same `key(item)` stay on one rank, in order."""
    if key is None:
        return [it for i, it in enumerate(items) if i % world_size == rank]
    groups = {}
    for it in items:
        groups.setdefault(key(it), []).append(it)
    mine = []
    for gi, k in enumerate(sorted(groups, key=str)):
        if gi % world_size == rank:
            mine.extend(groups[k])
    return mine


def scatter_items(items_on_rank0, key=None):
    """Rank 0 owns the item list (the driver reads it from disk); every rank gets its share."""
    rank, ws = world()
    if ws == 1:
        return list(items_on_rank0)
    shares = [assign_items(items_on_rank0, r, ws, key) for r in range(ws)] if rank == 0 else None
    out = [None]
    dist.scatter_object_list(out, shares, src=0)
    return out[0]


def frames_to_uint8(frames: torch.Tensor) -> torch.Tensor:
    """[-1,1] float frames -> uint8 with the reference's arithmetic (virtual_pose_render.py:243 clamp, eval_tools.py:24-27
    `(x + 1) / 2 * 255` then truncation).  A decoded clip on the GPU ([B, 3, T, H, W]) goes through the post-decode kernel
    (mudg_postdecode) and comes back as [B, T, 3, H, W]; anything else (host tensors in the plumbing tests) uses the same
    fp32 formula in torch."""
    if frames.is_cuda and frames.dim() == 5 and frames.shape[1] == 3:
        from .engine import MUDG_POST_COLOR, postdecode
        return postdecode(frames, [MUDG_POST_COLOR] * frames.shape[0])[0]
    return ((frames.float().clamp(-1.0, 1.0) + 1.0) / 2.0 * 255).to(torch.uint8)


def gather_frames(frames_u8: torch.Tensor, ids: torch.Tensor):
    """All-gather variable numbers of clips per rank: returns (frames [n_total, ...], ids [n_total]) sorted by id on
    every rank.  28 MB per 16x576x1024 clip -- off the critical path."""
    rank, ws = world()
    if ws == 1:
        order = torch.argsort(ids)
        return frames_u8[order], ids[order]
    n = torch.tensor([frames_u8.shape[0]], device=frames_u8.device)
    counts = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(counts, n)
    nmax = int(max(int(c) for c in counts))
    pad = torch.zeros((nmax,) + tuple(frames_u8.shape[1:]), dtype=frames_u8.dtype, device=frames_u8.device)
    pad[: frames_u8.shape[0]] = frames_u8
    pid = torch.full((nmax,), -1, dtype=torch.long, device=frames_u8.device)
    pid[: ids.shape[0]] = ids
    all_f = [torch.empty_like(pad) for _ in range(ws)]
    all_i = [torch.empty_like(pid) for _ in range(ws)]
    dist.all_gather(all_f, pad)
    dist.all_gather(all_i, pid)
    f = torch.cat(all_f)
    i = torch.cat(all_i)
    keep = i >= 0
    f, i = f[keep], i[keep]
    order = torch.argsort(i)
    return f[order], i[order]


def gather_frames_to(frames_u8: torch.Tensor, ids: torch.Tensor, dst: int = 0):
    """Gather the SAME number of clips from every rank onto rank `dst` only (the rank that writes the files,
    virtual_pose_render.py:243-274): returns (frames [n_total, ...], ids [n_total]) sorted by id on `dst`, (None, None)
    elsewhere.  One NCCL gather per call; 28 MB per 16x576x1024 uint8 clip and rank."""
    rank, ws = world()
    if ws == 1:
        order = torch.argsort(ids)
        return frames_u8[order], ids[order]
    frames_u8, ids = frames_u8.contiguous(), ids.contiguous()
    if rank == dst:
        fb = [torch.empty_like(frames_u8) for _ in range(ws)]
        ib = [torch.empty_like(ids) for _ in range(ws)]
        dist.gather(frames_u8, fb, dst=dst)
        dist.gather(ids, ib, dst=dst)
        f, i = torch.cat(fb), torch.cat(ib)
        order = torch.argsort(i)
        return f[order], i[order]
    dist.gather(frames_u8, None, dst=dst)
    dist.gather(ids, None, dst=dst)
    return None, None


def max_over_ranks(value: float, device=None) -> float:
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
