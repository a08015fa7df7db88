"""Multi-GPU = replicas only: clips shard over ranks, weights are broadcast once, the hot loop
has no collective.  One process per GPU (torch.distributed, NCCL over NVLink on the box, gloo in CPU tests)."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_clips(n_clips: int, world_size: int, rank: int) -> List[int]:
    """Clip i -> rank i mod world_size (config 5: 8 clips, one per GPU)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_clips, world_size))


def broadcast_state_dict(module: torch.nn.Module, src: int = 0, bucket_bytes: int = 64 << 20) -> int:
    """Broadcast every parameter/buffer of `module` from rank `src`, flattened into a few large buckets
    (sized for launch latency, not link count).  Returns the number of bytes sent.  In-place copies bump the
    parameter versions, so the model re-packs its device weights on the next forward."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    tensors = [t for _, t in sorted(module.state_dict().items())]
    total, bucket, size = 0, [], 0

    def flush():
        nonlocal bucket, size, total
        if not bucket:
            return
        flat = torch.cat([t.reshape(-1).float() for t in bucket])
        dist.broadcast(flat, src=src)
        off = 0
        with torch.no_grad():
            for t in bucket:
                n = t.numel()
                t.copy_(flat[off:off + n].reshape(t.shape).to(t.dtype))
                off += n
        total += flat.numel() * 4
        bucket, size = [], 0

    for t in tensors:
        bucket.append(t)
        size += t.numel() * 4
        if size >= bucket_bytes:
            flush()
    flush()
    return total


def gather_results(tracks: torch.Tensor, visibility: torch.Tensor, dst: int = 0):
    """Collect per-rank (tracks, visibility) on rank `dst` (≈1.2 MB per clip at N=6400, T=16)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [tracks], [visibility]
    world = dist.get_world_size()
    me = dist.get_rank()
    tl = [torch.empty_like(tracks) for _ in range(world)] if me == dst else None
    vl = [torch.empty_like(visibility, dtype=torch.uint8) for _ in range(world)] if me == dst else None
    dist.gather(tracks.contiguous(), tl, dst=dst)
    dist.gather(visibility.to(torch.uint8).contiguous(), vl, dst=dst)
    if me != dst:
        return None, None
    return tl, [v.bool() for v in vl]
