"""CPU, world_size 2 over gloo: the N>1 host logic -- contiguous sharding, the init-time
weight broadcast and the optional logits all-gather (SURVEY.md section 8 e)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import tskd_b200
from tskd_b200.dist import broadcast_weights, gather_logits, shard_range


def test_shard_range_partitions():
    for total in (1, 7, 4096, 32768, 4097):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(32768, 3, 8) == (12288, 16384)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)            # different weights per rank before the broadcast
        m = tskd_b200.B200MyCNN(tskd_b200.ARCH_PRESETS["mycnn5"].with_shape(3, 1500))
        before = m.packed_weights().clone()
        broadcast_weights(m, src=0)
        after = m.packed_weights()
        gathered = [torch.empty_like(after) for _ in range(world)]
        dist.all_gather(gathered, after)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        changed = (rank == 0) == bool(torch.equal(before, after))
        total = 11
        b, e = shard_range(total, rank, world)
        local = torch.arange(b, e, dtype=torch.float32) * 2
        full = gather_logits(local, total)
        ok = torch.equal(full, torch.arange(total, dtype=torch.float32) * 2)
        q.put((rank, same, changed, ok))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_gather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    for rank, same, changed, ok in res:
        assert same and changed and ok, (rank, same, changed, ok)
