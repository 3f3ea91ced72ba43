"""The N>1 path of bench.py on CPU: two gloo ranks run the ring's control plane (descriptor exchange,
predecessor selection, lock-step hops) with fake 256-byte descriptors; the payload movers are CUDA-only
and are covered by the `-m gpu` tests and the N=2 bench."""
import os
import socket
import sys
from pathlib import Path

import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_ops: int, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lzy_b200.ring import DESC_BYTES, exchange_descriptors, hop_plan, predecessor
    plan = hop_plan(world, n_ops)
    pulled = []
    token = rank  # what this rank's chain currently holds: the id of the chain that started on `token`
    for h in range(n_ops):
        desc = bytes([rank, h, token]) + bytes(DESC_BYTES - 3)
        descs = exchange_descriptors(dist, desc)
        src = plan[h][rank]
        assert src == predecessor(rank, world)
        r, hop, tok = descs[src][0], descs[src][1], descs[src][2]
        assert (r, hop) == (src, h)
        token = tok  # "pull": the chain moves one step round the ring
        pulled.append(src)
        dist.barrier()
    out.put((rank, token, pulled))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_ring_control_plane_two_ranks(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_ops = 8
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_ops, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, token, pulled in res:
        assert pulled == [(rank - 1) % world] * n_ops
        assert token == (rank - n_ops) % world  # after 8 hops the chain that started 8 ranks upstream is here


def _mailbox_worker(rank: int, world: int, port: int, n_ops: int, steps: int, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lzy_b200.ring import DESC_BYTES, ShmMailbox, predecessor, successor
    box = ShmMailbox.rendezvous(dist, world, rank, lanes=2)
    SINK, OUT = 0, 1
    token = rank
    try:
        for step in range(steps):
            token = rank  # every step starts a fresh chain on every rank
            for h in range(n_ops):
                # what bench.py's push ring does per hop: publish my sink and my committed output, learn my successor's
                # sink (my push kernel writes there) and my predecessor's output (what my scan verifies against)
                box.post(SINK, bytes([rank, step, h, 0xAA]) + bytes(DESC_BYTES - 4))
                box.post(OUT, bytes([rank, step, h, token]) + bytes(DESC_BYTES - 4))
                sink = box.read(successor(rank, world), SINK)
                src = box.read(predecessor(rank, world), OUT)
                assert tuple(sink[:4]) == (successor(rank, world), step, h, 0xAA)
                assert tuple(src[:3]) == (predecessor(rank, world), step, h)
                token = src[3]
            dist.barrier()
        out.put((rank, token))
    finally:
        box.close()
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_control_plane_over_the_shared_memory_mailbox(world):
    """descriptor exchange without a collective: posts and reads pair up one to one per (rank, lane), across steps"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_ops, steps = 8, 5
    procs = [ctx.Process(target=_mailbox_worker, args=(r, world, port, n_ops, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, token in res:
        assert token == (rank - n_ops) % world


def test_mailbox_refuses_wrong_sizes_and_times_out():
    from lzy_b200.ring import ShmMailbox
    import uuid
    box = ShmMailbox(f"lzyb_t_{uuid.uuid4().hex[:10]}", 2, 0, lanes=1, create=True)
    try:
        with pytest.raises(ValueError):
            box.post(0, b"short")
        with pytest.raises(TimeoutError):
            box.read(1, 0, timeout=0.05)
    finally:
        box.close()


def test_descriptor_size_is_checked():
    from lzy_b200.ring import exchange_descriptors
    with pytest.raises(ValueError):
        exchange_descriptors(None, b"short")


def _leaver(rank: int, world: int, port: int, out):
    """bench.py's sequence around the one-process legs (run_ours): the ranks learn each other's pids over a gloo HOST group,
    meet at a host barrier, then every rank but 0 tears its process group down (bounded) and leaves while rank 0 waits for
    their processes to be gone and goes on alone — and can still destroy its own group afterwards"""
    import threading
    import time
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    host_group = dist.new_group(backend="gloo")
    pids = [None] * world
    dist.all_gather_object(pids, os.getpid(), group=host_group)
    dist.barrier(group=host_group)
    if rank != 0:
        t = threading.Thread(target=dist.destroy_process_group, daemon=True)
        t.start()
        t.join(5.0)
        out.put((rank, pids, None))
        if t.is_alive():
            os._exit(0)
        return
    import bench
    t0 = time.time()
    waited = bench.wait_until_alone([p for r, p in enumerate(pids) if r != rank], timeout_s=30.0)
    alone = all(not os.path.exists(f"/proc/{p}") or open(f"/proc/{p}/stat").read().rsplit(")", 1)[1].split()[0] in "ZX"
                for r, p in enumerate(pids) if r != rank)
    dist.destroy_process_group()  # the peers are gone: must not hang
    out.put((rank, pids, (waited, alone, time.time() - t0)))


def test_other_ranks_leave_and_rank_0_goes_on_alone():
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_leaver, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {r: (pids, extra) for r, pids, extra in (q.get(timeout=120) for _ in range(world))}
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(res[r][0] == [p.pid for p in procs] for r in range(world))  # everybody learnt everybody's pid, in rank order
    waited, alone, total = res[0][1]
    assert alone and waited < 25.0 and total < 40.0
