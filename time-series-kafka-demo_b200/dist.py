"""Multi-GPU layout for the path: one process per GPU, the window batch sharded contiguously
("one patient-stream per shard"), no data-path collective.

Independent windows shard trivially (SURVEY.md section 8 e): the only communication is the
init-time broadcast of the small packed weight blob (<= 4.8 MB) from rank 0 -- NCCL over
NVLink on the GPU box, gloo in the CPU tests -- and an optional all-gather of the B logits.
The reference has no distributed mode at all (Spark ``local[1]``, bin/utils.py:94).
``sequence`` mode (LSTM scanning the batch axis) does not shard: replicas only.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of `total` windows for `rank` (first ranks take the
    remainder)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def broadcast_weights(model, src: int = 0, group=None) -> None:
    """Rank `src`'s parameters overwrite every rank's: ONE broadcast of the packed blob plus
    the inert tensors, then load_state_dict on the receivers."""
    sd = model.state_dict()
    keys = sorted(sd.keys())
    flat = torch.cat([sd[k].detach().reshape(-1).float() for k in keys]).contiguous()
    dist.broadcast(flat, src=src, group=group)
    if dist.get_rank(group) != src:
        off = 0
        new = {}
        for k in keys:
            n = sd[k].numel()
            new[k] = flat[off:off + n].view_as(sd[k]).to(sd[k].dtype)
            off += n
        model.load_state_dict(new)


def gather_logits(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather the per-shard logits into the full [total] vector (optional: B*4 bytes)."""
    world = dist.get_world_size(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    maxn = max(e - b for b, e in sizes)
    pad = torch.zeros(maxn, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][:e - b] for r, (b, e) in enumerate(sizes)])
