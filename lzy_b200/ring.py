"""Control plane of the N-worker ring used by ``bench.py`` when it is launched one process per GPU.

Only slot descriptors cross this path — never payload bytes.  At every hop each worker publishes

* the descriptor of the output slot it just committed (digest, dtype, shape: what its consumer verifies against), and
* the descriptor of its own empty input frame (the *sink* its predecessor's push kernel writes into),

and learns its predecessor's output descriptor and its successor's sink descriptor.

Round 1 did this with an NCCL all-gather of the descriptor plus ``.cpu()`` per hop — ≈ 0.4 ms of host time per hop,
18 % of the 8-GPU step.  ``ShmMailbox`` replaces it with a POSIX shared-memory board: a worker
writes its 256-byte descriptor and then bumps a sequence word; its neighbour spins on that word.  A post or a read is a
memcpy of 256 bytes — microseconds — and involves neither the GPU nor a collective.  ``exchange_descriptors`` (the
collective version) is kept for the CPU tests that run the ring logic over gloo.

What this stands in for in the reference: the channel manager telling a consumer who its producer is
(``Bind`` → ``StartTransfer``, channel-manager/.../services/SlotsService.java:98-145,440-519) — here both ends are
known from the ring position, so the "control RPC" is a mailbox read.
"""
from __future__ import annotations

import struct
import time
from multiprocessing import shared_memory
from typing import List, Optional

DESC_BYTES = 256
_ENTRY = 8 + DESC_BYTES  # sequence word + descriptor
_DEPTH = 4               # cells per (rank, lane): a writer may run this many posts ahead of a reader minus one


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


class ShmMailbox:
    """A board of ``world × lanes`` descriptor cells in POSIX shared memory.

    ``post(lane, desc)`` publishes this rank's descriptor on a lane; ``read(rank, lane)`` blocks until that rank's
    next post on that lane and returns it.  Every (rank, lane) cell carries a sequence number that only its owner
    advances; a reader remembers the last sequence it consumed per cell, so posts and reads pair up one to one in
    order without any other synchronisation.  Post #s of a lane lands in cell s mod 4, so a writer may run three
    posts ahead of its slowest reader; in the ring a worker is never more than one hop ahead of its successor (its next
    push needs the successor's next sink descriptor).
    """

    def __init__(self, name: str, world: int, rank: int, lanes: int = 2, create: bool = False):
        self.world, self.rank, self.lanes = world, rank, lanes
        size = world * lanes * _DEPTH * _ENTRY
        self._shm = shared_memory.SharedMemory(name=name, create=create, size=size if create else 0)
        if create:
            self._shm.buf[:size] = bytes(size)
        else:
            # attaching registers the segment with this process's resource tracker as if it were ours (bpo-39959): it would
            # unlink the creator's segment — or complain that it is gone — when this process exits
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:
                pass
        self._buf = self._shm.buf
        self._posted = [0] * lanes                            # sequence of my last post per lane
        self._seen = [[0] * lanes for _ in range(world)]      # last sequence consumed per (rank, lane)
        self._owner = create
        self.name = name

    @classmethod
    def rendezvous(cls, dist, world: int, rank: int, lanes: int = 2) -> "ShmMailbox":
        """rank 0 creates the board, everybody learns its name through one object broadcast (set-up time only)"""
        import uuid
        box: Optional[ShmMailbox] = None
        names = [None]
        if rank == 0:
            box = cls(f"lzyb_ring_{uuid.uuid4().hex[:12]}", world, rank, lanes, create=True)
            names[0] = box.name
        dist.broadcast_object_list(names, src=0)
        if rank != 0:
            box = cls(names[0], world, rank, lanes, create=False)
        dist.barrier()
        return box

    def _off(self, rank: int, lane: int, seq: int) -> int:
        return ((rank * self.lanes + lane) * _DEPTH + seq % _DEPTH) * _ENTRY

    def post(self, lane: int, desc: bytes) -> None:
        if len(desc) != DESC_BYTES:
            raise ValueError(f"a slot descriptor is {DESC_BYTES} bytes, got {len(desc)}")
        self._posted[lane] += 1
        off = self._off(self.rank, lane, self._posted[lane])
        self._buf[off + 8: off + 8 + DESC_BYTES] = desc
        struct.pack_into("<Q", self._buf, off, self._posted[lane])  # the sequence word goes last (x86: stores stay in order)

    def read(self, rank: int, lane: int, timeout: float = 60.0) -> bytes:
        want = self._seen[rank][lane] + 1
        off = self._off(rank, lane, want)
        deadline = None
        spins = 0
        while True:
            (seq,) = struct.unpack_from("<Q", self._buf, off)
            if seq >= want:
                break
            spins += 1
            if spins > 2000:  # ≈ 1 ms of pure spinning first: the neighbour is normally microseconds away
                if deadline is None:
                    deadline = time.monotonic() + timeout
                elif time.monotonic() > deadline:
                    raise TimeoutError(f"rank {self.rank}: no descriptor #{want} from rank {rank} on lane {lane}")
                time.sleep(0)
        if seq != want:
            raise RuntimeError(f"rank {self.rank}: descriptor #{want} from rank {rank} lane {lane} was overwritten (now #{seq})")
        data = bytes(self._buf[off + 8: off + 8 + DESC_BYTES])
        (again,) = struct.unpack_from("<Q", self._buf, off)
        if again != want:
            raise RuntimeError(f"rank {self.rank}: descriptor #{want} from rank {rank} lane {lane} changed while it was read")
        self._seen[rank][lane] = want
        return data

    def close(self) -> None:
        self._buf = None
        try:
            self._shm.close()
            if self._owner:
                self._shm.unlink()
        except Exception:
            pass
