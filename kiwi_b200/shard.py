"""Sentence sharding across GPUs (BASELINE.json config 5: round-robin by sentence index, i -> GPU i mod N).
The path has no exchange step: the read-only model image is broadcast once at init, every rank analyses its
shard, the launcher restores input order.  (The reference's only parallelism is its ordered thread pool,
include/kiwi/Kiwi.h:402-454.)"""
from typing import List, Sequence


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n, world))


def merge_round_robin(per_rank: Sequence[Sequence], n: int) -> list:
    world = len(per_rank)
    out = [None] * n
    for r, items in enumerate(per_rank):
        for k, item in enumerate(items):
            out[r + k * world] = item
    return out


def broadcast_image(image_bytes, dist, device=None):
    """One collective at init: rank 0's model image to every rank (NCCL on GPU tensors, gloo on CPU tensors)."""
    import torch
    rank = dist.get_rank()
    size = torch.tensor([len(image_bytes) if rank == 0 else 0], dtype=torch.int64, device=device)
    dist.broadcast(size, 0)
    if rank == 0:
        buf = torch.frombuffer(bytearray(image_bytes), dtype=torch.uint8).to(device) if device else torch.frombuffer(bytearray(image_bytes), dtype=torch.uint8)
    else:
        buf = torch.empty(int(size.item()), dtype=torch.uint8, device=device)
    dist.broadcast(buf, 0)
    return buf.cpu().numpy().tobytes()
