"""Host memory next to the GPU that will read it.

A B200 box has two sockets and every GPU hangs off one of them; a pinned staging buffer on the other socket makes each
H2D / D2H copy cross the inter-socket link.  `cudaHostAlloc` places pages on the node of the calling thread, so the
placement is decided by where the allocating thread runs: `near(gpu)` moves it next to the GPU for the duration of the
allocation, `pinned_empty` is the allocation itself.  Everything here is best effort — in a container whose cpuset forbids
the move the allocation simply stays where it was."""
from __future__ import annotations

import contextlib
import os
from typing import Iterator, Optional, Set


def _parse_cpulist(text: str) -> Set[int]:
    cpus: Set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def local_cpus(gpu: int) -> Optional[Set[int]]:
    """CPUs of the socket the GPU is attached to (sysfs `local_cpulist` of its PCI function), or None if unknown."""
    try:
        import torch

        p = torch.cuda.get_device_properties(gpu)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            cpus = _parse_cpulist(f.read())
        return cpus or None
    except Exception:  # noqa: BLE001 — no sysfs, no such attribute, no GPU: placement stays as it is
        return None


@contextlib.contextmanager
def near(gpu: int) -> Iterator[bool]:
    """Run the calling thread on the GPU's socket inside the block; yields whether the move happened."""
    cpus = local_cpus(gpu)
    moved = False
    before = None
    if cpus and hasattr(os, "sched_setaffinity"):
        try:
            before = os.sched_getaffinity(0)
            target = cpus & before or cpus
            os.sched_setaffinity(0, target)
            moved = True
        except OSError:
            moved = False
    try:
        yield moved
    finally:
        if moved and before:
            try:
                os.sched_setaffinity(0, before)
            except OSError:
                pass


def bind_process(gpu: int) -> bool:
    """One process per GPU: keep the whole process (its threads-to-be and its allocations) on the GPU's socket."""
    cpus = local_cpus(gpu)
    if not cpus or not hasattr(os, "sched_setaffinity"):
        return False
    try:
        allowed = os.sched_getaffinity(0)
        os.sched_setaffinity(0, cpus & allowed or cpus)
        return True
    except OSError:
        return False


def pinned_empty(shape, dtype, gpu: int):
    """`torch.empty(..., pin_memory=True)` with its pages on the GPU's socket (first touched there as well)."""
    import torch

    with near(gpu):
        t = torch.empty(shape, dtype=dtype, pin_memory=True)
        t.view(torch.uint8).reshape(-1)[:: 4096].zero_()
    return t
