"""Multi-GPU: independent clips shard data-parallel, one process per GPU ("replicas only").

Frame i of a clip consumes stylized frame i-1 (fast_artistic_video.lua:153-158,168-169), so nothing inside a clip
parallelises over time; clips share only read-only weights.  The only communication is moving inputs from the rank
that decoded them to the owning rank and collecting outputs -- torch.distributed point-to-point (NCCL over NVLink on
the GPU box, gloo in the CPU tests).  No collective sits on the data path of a frame.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


def assign_clips(num_clips: int, world_size: int) -> List[List[int]]:
    """Round-robin: clip c is owned by rank c % world_size."""
    return [[c for c in range(num_clips) if c % world_size == r] for r in range(world_size)]


def owner(clip: int, world_size: int) -> int:
    return clip % world_size


def scatter_clips(clips: Optional[List[torch.Tensor]], shapes: List[tuple], src: int = 0, device=None,
                  dtype=torch.float32) -> Dict[int, torch.Tensor]:
    """Rank `src` holds clips[c] (any tensor per clip, e.g. [T,3,H,W] frames); every rank returns {clip: tensor} for the
    clips it owns.  shapes[c] is known on all ranks (from the container headers)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    mine: Dict[int, torch.Tensor] = {}
    reqs = []
    for c, shp in enumerate(shapes):
        o = owner(c, world)
        if rank == src:
            if o == src:
                mine[c] = clips[c].to(device) if device is not None else clips[c]
            else:
                t = clips[c].to(device) if device is not None else clips[c]
                reqs.append(dist.isend(t.contiguous(), dst=o, tag=c))
        elif rank == o:
            buf = torch.empty(shp, dtype=dtype, device=device)
            reqs.append(dist.irecv(buf, src=src, tag=c))
            mine[c] = buf
    for r in reqs:
        r.wait()
    return mine


def gather_clips(local: Dict[int, torch.Tensor], shapes: List[tuple], dst: int = 0, device=None,
                 dtype=torch.float32) -> Optional[List[torch.Tensor]]:
    """Inverse of scatter_clips: rank `dst` returns the list of all clips' outputs, other ranks None."""
    rank, world = dist.get_rank(), dist.get_world_size()
    out: List[Optional[torch.Tensor]] = [None] * len(shapes)
    reqs = []
    for c, shp in enumerate(shapes):
        o = owner(c, world)
        if rank == dst:
            if o == dst:
                out[c] = local[c]
            else:
                buf = torch.empty(shp, dtype=dtype, device=device)
                reqs.append(dist.irecv(buf, src=o, tag=c))
                out[c] = buf
        elif rank == o:
            reqs.append(dist.isend(local[c].contiguous(), dst=dst, tag=c))
    for r in reqs:
        r.wait()
    return out if rank == dst else None
