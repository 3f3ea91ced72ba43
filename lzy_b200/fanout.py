"""Fan-out of one output slot to many consumers in *different worker processes* (one per GPU):
a single ``ncclBroadcast`` over NVSwitch (NVLS multicast where NCCL enables it) straight into the
consumers' input-slot frames, then every consumer verifies what landed against the producer's digest
with the scan kernel.  This is the one exchange step of the path: the reference serves
R consumers with R independent unicast ``Read`` streams (channel-manager/.../services/SlotsService.java:457-466,
lzy/slots/.../OutputSlot.java:66-72).  Inside one process the chunk-pipelined pull chain
(``lzyb_broadcast``) does the same job without a communicator."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _native as N
from .devmem import view_bytes

_META_WORDS = 2 + 1 + 1 + N.MAX_DIMS + 1  # digest[2], dtype, ndim, shape[8], nbytes


def broadcast_slot_nccl(dist, root: int, src: Optional[N.SlotHandle], gpu: int, slot_id: str,
                        stream: Optional[torch.cuda.Stream] = None) -> Tuple[N.SlotHandle, float]:
    """Collective over the default process group.  On `root`, `src` is the committed output slot; every
    other rank passes None and gets back a READY input slot holding the same bytes.  Returns
    (slot, milliseconds spent in the broadcast itself, CUDA-event timed on this rank)."""
    rank = dist.get_rank()
    dev = f"cuda:{gpu}"
    meta = torch.zeros(_META_WORDS, dtype=torch.int64, device=dev)
    if rank == root:
        info = src.info()
        vals = [int(info.digest[0]), int(info.digest[1]), int(info.dtype), int(info.ndim)] + \
               [int(info.shape[i]) for i in range(N.MAX_DIMS)] + [int(info.nbytes)]
        meta.copy_(torch.tensor([v - (1 << 64) if v >= (1 << 63) else v for v in vals], dtype=torch.int64))
    dist.broadcast(meta, src=root)
    vals = [int(v) & ((1 << 64) - 1) for v in meta.tolist()]
    digest, dtype, ndim, shape, nbytes = (vals[0], vals[1]), vals[2], vals[3], vals[4:4 + N.MAX_DIMS], vals[-1]
    if rank == root:
        slot = src
    else:
        slot = N.SlotHandle.open_input(slot_id, gpu, nbytes)
    payload = view_bytes(slot.device_ptr(), nbytes, gpu)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = stream or torch.cuda.current_stream(gpu)
    with torch.cuda.stream(st):
        e0.record(st)
        dist.broadcast(payload, src=root)
        e1.record(st)
    e1.synchronize()
    if rank != root:
        slot.adopt_external(digest, dtype, shape[:ndim], int(st.cuda_stream) or 1)
    return slot, e0.elapsed_time(e1)
