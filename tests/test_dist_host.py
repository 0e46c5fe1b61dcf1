"""Host-side multi-GPU logic on CPU: partition math and the unique-id rendezvous over a world_size-2 gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import product

product()
from ltx_video_swift_mlx_b200 import dist as ltxdist  # noqa: E402


@pytest.mark.parametrize("frames,world", [(4, 1), (4, 2), (4, 4), (4, 8), (16, 8), (16, 3), (33, 8), (1, 8)])
def test_temporal_slabs_cover_all_frames_once(frames, world):
    slabs = ltxdist.temporal_slabs(frames, world)
    assert len(slabs) == min(world, frames)
    assert slabs[0][0] == 0 and slabs[-1][1] == frames
    assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:])) and all(f1 > f0 for f0, f1 in slabs)
    sizes = [f1 - f0 for f0, f1 in slabs]
    assert max(sizes) - min(sizes) <= 1
    outs = [ltxdist.slab_output_frames(*s) for s in slabs]
    assert outs[0][0] == 0 and outs[-1][1] == 8 * (frames - 1) + 1            # V/VideoDecoder.swift:294 frame formula
    assert all(a[1] == b[0] for a, b in zip(outs, outs[1:]))


def test_pass_and_token_partition():
    for groups in (1, 2, 3, 4):
        owners = [ltxdist.pass_owner(p, groups) for p in range(3)]
        assert all(0 <= o < groups for o in owners) and owners[0] == 0
    assert [ltxdist.token_shard(1536, 8, r) for r in (0, 7)] == [(0, 192), (1344, 1536)]
    with pytest.raises(ValueError):
        ltxdist.token_shard(1537, 8, 0)
    assert [ltxdist.rank_layout(r, 4) for r in (0, 3, 4, 7)] == [(0, 0), (0, 3), (1, 0), (1, 3)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = ltxdist.exchange_unique_id(lambda: bytes(range(128)), rank)
        # every rank derives the same partition and the union is exact
        slabs = ltxdist.temporal_slabs(5, world)
        mine = torch.tensor(slabs[rank] if rank < len(slabs) else (0, 0))
        allr = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
        dist.all_gather(allr, mine)
        q.put((rank, uid == bytes(range(128)), [tuple(t.tolist()) for t in allr] == slabs))
    finally:
        dist.destroy_process_group()


def test_unique_id_rendezvous_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] and r[2] for r in res)
