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


def test_descriptor_size_is_checked():
    from lzy_b200.ring import exchange_descriptors
    with pytest.raises(ValueError):
        exchange_descriptors(None, b"short")
