"""Multi-GPU host logic: one process per GPU (torchrun), `torch.distributed` for the rendezvous only; the data-path
collectives (NCCL all-to-all / send-recv / broadcast) are issued by libltxcuda on its own communicators.

The reference has no multi-device code: the partitioning below is new design; the contract is
N-GPU result == 1-GPU result."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


def temporal_slabs(latent_frames: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous latent-frame slabs [f0, f1) of the ranks that take part in a sharded VAE decode (the first
    min(world, F') ranks; balanced, earlier ranks get the remainder).  Mirrors vae.cu::slab."""
    n = min(world_size, latent_frames)
    base, rem = divmod(latent_frames, n)
    out, f0 = [], 0
    for r in range(n):
        f1 = f0 + base + (1 if r < rem else 0)
        out.append((f0, f1))
        f0 = f1
    return out


def slab_output_frames(f0: int, f1: int) -> Tuple[int, int]:
    """Decoded frame range of a latent slab: latent frame f produces frames 8(f-1)+1 .. 8f (frame 0 for f = 0)."""
    return (0 if f0 == 0 else 8 * (f0 - 1) + 1, 8 * (f1 - 1) + 1)


def pass_owner(pass_index: int, pass_groups: int) -> int:
    """Group that runs guidance pass `pass_index` (0 conditional, 1 unconditional, 2 STG; absent passes are skipped
    before numbering).  Mirrors ltx_denoise_step."""
    return pass_index % pass_groups


def token_shard(num_tokens: int, sp_size: int, sp_rank: int) -> Tuple[int, int]:
    """Token range of an Ulysses rank (N must be divisible by sp_size)."""
    if num_tokens % sp_size:
        raise ValueError("token count must be divisible by the sequence-parallel degree")
    n = num_tokens // sp_size
    return sp_rank * n, (sp_rank + 1) * n


def rank_layout(rank: int, sp_size: int) -> Tuple[int, int]:
    """(pass group, sp rank) of a world rank: rank = group * sp_size + sp_rank."""
    return rank // sp_size, rank % sp_size


def exchange_unique_id(make_id, rank: int) -> bytes:
    """Rank 0 creates the 128-byte NCCL id (make_id()), every rank returns it.  Works on any torch.distributed backend."""
    import torch.distributed as dist
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("bad NCCL unique id")
    return bytes(uid)


def init_context(ctx, sp_size: int = 1, pass_groups: Optional[int] = None):
    """Initialises the library communicators of `ctx` from an already initialised torch.distributed process group."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if pass_groups is None:
        pass_groups = world // sp_size
    uid = exchange_unique_id(type(ctx).dist_unique_id, rank)
    ctx.dist_init(uid, rank, world, sp_size, pass_groups)
    return rank, world
