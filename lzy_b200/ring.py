"""Control plane of the N-worker ring used by ``bench.py`` when it is launched one process per GPU:
at every hop each worker publishes the 256-byte descriptor of the slot it just committed and learns
its predecessor's.  Only descriptors cross this path (``torch.distributed`` all-gather, NCCL on GPUs,
gloo in the CPU tests); payload bytes move by the CUDA pull kernel through the CUDA-IPC mapping."""
from __future__ import annotations

from typing import List

DESC_BYTES = 256


def predecessor(rank: int, world: int) -> int:
    return (rank - 1) % world


def successor(rank: int, world: int) -> int:
    return (rank + 1) % world


def exchange_descriptors(dist, desc: bytes, device: str = "cpu") -> List[bytes]:
    """all-gather one descriptor per rank; returns the list indexed by rank"""
    import torch
    if len(desc) != DESC_BYTES:
        raise ValueError(f"a slot descriptor is {DESC_BYTES} bytes, got {len(desc)}")
    mine = torch.frombuffer(bytearray(desc), dtype=torch.uint8).to(device)
    world = dist.get_world_size()
    out = [torch.zeros(DESC_BYTES, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


def hop_plan(world: int, n_ops: int) -> List[List[int]]:
    """plan[h][r] = rank whose slot rank r pulls at hop h (the N chains are staggered round the ring,
    so every rank pulls from its predecessor at every hop and every link carries one payload per hop)"""
    return [[predecessor(r, world) for r in range(world)] for _ in range(n_ops)]
