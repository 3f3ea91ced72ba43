"""Fan-out of one output slot to R ≥ 3 consumers that live in OTHER worker processes (one per GPU): one `ncclBroadcast`
over NVSwitch straight into the consumers' input-slot frames, every consumer verifying what landed against the
producer's digest with the scan kernel (lzy_b200/fanout.py; SURVEY §8e — the path's one exchange step).  The reference
serves R consumers with R independent Read streams (channel-manager/.../services/SlotsService.java:457-466,
lzy/slots/.../OutputSlot.java:66-72); SlotsIntegrationTest.testMultiConsumer (:212-297) is the scenario: 1 → 3, all equal.
Bit-exact against the counter-based pattern of the oracle.  Needs ≥ 4 GPUs (`gpurun --gpus 4`)."""
import os
import socket
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available() or torch.cuda.device_count() < 4:
    pytest.skip("needs at least 4 GPUs", allow_module_level=True)

import torch.multiprocessing as mp  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent
N_BYTES = (64 << 20) + 5
SEED = 7


def _worker(rank: int, world: int, port: int, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import numpy as np
    import torch
    import torch.distributed as dist

    import oracle
    from lzy_b200 import _native as N
    from lzy_b200.devmem import view_bytes
    from lzy_b200.fanout import broadcast_slot_nccl
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    N.init(rank, 1, 16 << 20)
    try:
        src = None
        if rank == 0:
            t = torch.empty(N_BYTES, dtype=torch.uint8, device="cuda:0")
            N.fill_pattern(0, t.data_ptr(), N_BYTES, SEED)
            torch.cuda.synchronize()
            src = N.SlotHandle.create_output("fan/src", 0, N_BYTES)
            src.set_meta(N.DT_U8, (N_BYTES,))
            src.pack(t.data_ptr(), N_BYTES)
            src.commit()
        slot, _ms = broadcast_slot_nccl(dist, 0, src, rank, f"fan/in{rank}")
        expect = oracle.pattern(N_BYTES, SEED)
        info = slot.info()
        ok = (info.state == N.ST_READY and tuple(info.digest) == oracle.digest(expect) and int(info.nbytes) == N_BYTES
              and info.dtype == N.DT_U8 and np.array_equal(view_bytes(slot.device_ptr(), N_BYTES, rank).cpu().numpy(), expect))
        out.put((rank, bool(ok), bool(info.is_input)))
        dist.barrier()
        slot.close()
    finally:
        dist.destroy_process_group()


def test_one_nccl_broadcast_serves_every_consumer_process_bit_exact():
    world = min(torch.cuda.device_count(), 8)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert [r for r, _, _ in res] == list(range(world))
    assert all(ok for _, ok, _ in res)
    assert [is_in for _, _, is_in in res] == [False] + [True] * (world - 1)  # R = world − 1 ≥ 3 consumers, one producer
