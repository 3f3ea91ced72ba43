"""Multi-GPU behaviour inside one process (workers pinned one per GPU, peer access over NVSwitch):
cross-GPU consumer-pull, fan-out, and the runtime placing an @op chain / diamond across GPUs
(BASELINE configs C2, C3, C5 at test sizes).  Needs ≥ 2 GPUs: run with `gpurun --gpus 2`."""
import uuid
from typing import Dict

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs at least 2 GPUs", allow_module_level=True)

from lzy_b200 import _native as N  # noqa: E402
from lzy_b200.api.v1 import B200Runtime, Lzy, op  # noqa: E402

NGPU = torch.cuda.device_count()


@pytest.fixture(scope="module")
def native_all():
    N.init(0, 0, 64 << 20)
    return N


def to_host(slot):
    out = np.empty(slot.nbytes, dtype=np.uint8)
    slot.copy_to_host(out.ctypes.data, slot.nbytes)
    return out


@pytest.mark.parametrize("engine", ["ldg", "tma", "dma", "push", "auto"])
def test_cross_gpu_pull_is_bit_exact(native_all, engine):
    N.set_copy_engine({"ldg": N.ENGINE_LDG, "tma": N.ENGINE_TMA, "dma": N.ENGINE_DMA, "push": N.ENGINE_PUSH,
                       "auto": N.ENGINE_AUTO}[engine])
    N.set_dma_schedule(3 << 20, 1)  # several sub-chunks (and a ragged last one) per 8 MiB chunk
    try:
        n = (48 << 20) + 4096 + 7
        a = np.random.default_rng(5).integers(0, 256, n, dtype=np.uint8)
        t = torch.from_numpy(a).to("cuda:0")
        out = N.SlotHandle.create_output("x/out", 0, n)
        inp = N.SlotHandle.open_input("x/in", 1, n)
        try:
            out.pack(t.data_ptr(), n)
            dg = out.commit()
            before = N.stats()["bytes_pulled_peer"]
            off = 0
            while True:
                moved = inp.transfer_chunk(out, off, 8 << 20)
                if moved == N.EOF:
                    break
                off += moved
            inp.finish(out)
            assert N.stats()["bytes_pulled_peer"] - before == n
            assert inp.info().gpu == 1 and tuple(inp.info().digest) == dg == oracle.digest(a)
            assert np.array_equal(to_host(inp), a)
        finally:
            out.close()
            inp.close()
    finally:
        N.set_copy_engine(N.ENGINE_AUTO)
        N.set_dma_schedule()


def test_copy_engine_pull_trailing_scan_catches_corruption(native_all):
    """The copy engine moves the bytes, the trailing scans own the digest: a peer that serves one wrong byte
    is a ReadException at EOS exactly as with the fused kernels, and the sink can refetch elsewhere."""
    n = (40 << 20) + 11
    a = np.random.default_rng(6).integers(0, 256, n, dtype=np.uint8)
    t = torch.from_numpy(a).to("cuda:0")
    bad = N.SlotHandle.create_output("dma/bad", 0, n)
    good = N.SlotHandle.create_output("dma/good", 0, n)
    inp = N.SlotHandle.open_input("dma/in", 1, n)
    N.set_copy_engine(N.ENGINE_DMA)
    N.set_dma_schedule(4 << 20, 3)  # 40 MiB + 11 B = 24 + 12 + (4 MiB + 11 B)
    try:
        for s in (bad, good):
            s.pack(t.data_ptr(), n)
            s.commit()
        bad.fault(N.FAULT_CORRUPT_ONCE, (17 << 20) + 5)
        before = N.stats()["dma_copies"]
        assert inp.transfer_chunk(bad, 0, 1 << 40) == n
        assert N.stats()["dma_copies"] - before == 3
        with pytest.raises(N.ReadException):
            inp.finish(bad)
        assert inp.info().delivered == 0
        assert inp.transfer_chunk(good, 0, 1 << 40) == n
        inp.finish(good)
        assert tuple(inp.info().digest) == oracle.digest(a)
        assert np.array_equal(to_host(inp), a)
    finally:
        N.set_copy_engine(N.ENGINE_AUTO)
        N.set_dma_schedule()
        for s in (bad, good, inp):
            s.close()


@pytest.mark.parametrize("n,step", [((64 << 20) + 4096 + 7, 0), ((8 << 20) + 16, 0), (32768 * 3 + 5, 0), ((40 << 20) + 48, (8 << 20) + 16),
                                    ((1 << 20) + 1, 0), (1 << 30, 0)],
                         ids=["64MiB-ragged", "8MiB+16", "3tiles+5", "chunked-16B-offsets", "1MiB+1", "1GiB"])
def test_push_from_the_source_gpu_delivers_bit_exact(native_all, n, step):
    """The default carrier between two GPUs of one process: the streaming kernel launched on the SOURCE GPU, storing into
    the sink frame through the peer mapping (TMA ring, digest folded from shared memory; the LDG kernel for short or
    16-byte-offset chunks).  Byte equality against the source, the digest against the oracle, no copy engine; a second
    delivery reuses the pooled frame."""
    N.set_copy_engine(N.ENGINE_PUSH)
    try:
        if n >= (256 << 20):
            t = torch.empty(n, dtype=torch.uint8, device="cuda:0")
            N.fill_pattern(0, t.data_ptr(), n, 21)
            torch.cuda.synchronize(0)
            a = None
        else:
            a = np.random.default_rng(n & 0xFFFF).integers(0, 256, n, dtype=np.uint8)
            t = torch.from_numpy(a).to("cuda:0")
        out = N.SlotHandle.create_output("push/out", 0, n)
        out.pack(t.data_ptr(), n)
        dg = out.commit()
        if a is not None:
            assert dg == oracle.digest(a)
        for life in range(2):  # the second life reuses the pooled frame: its counters must be zero again
            inp = N.SlotHandle.open_input(f"push/in{life}", 1, n)
            before = N.stats()
            off = 0
            while True:
                moved = inp.transfer_chunk(out, off, step or (1 << 40))
                if moved == N.EOF:
                    break
                off += moved
            inp.finish(out)
            after = N.stats()
            assert after["dma_copies"] == before["dma_copies"]
            assert after["push_launches"] - before["push_launches"] == (1 if not step else -(-n // step))
            assert after["bytes_pulled_peer"] - before["bytes_pulled_peer"] == n
            assert inp.info().gpu == 1 and inp.info().state == N.ST_READY and tuple(inp.info().digest) == dg
            got = torch.empty(n, dtype=torch.uint8, device="cuda:1")
            from lzy_b200.devmem import view_bytes
            got.copy_(view_bytes(inp.device_ptr(), n, 1))
            assert torch.equal(got.to("cuda:0"), t)
            inp.close()
        out.close()
    finally:
        N.set_copy_engine(N.ENGINE_AUTO)


def test_push_catches_a_corrupted_source_and_the_sink_refetches(native_all):
    """a frame that changed after its commit is a peer fault: the push kernel's digest of what it hands to the wire differs
    from the committed one — a ReadException at EOS; the sink is reset and fetches from another producer
    (SlotsTest.testRestoreReadFromOffset's failover, with the new carrier)"""
    n = (24 << 20) + 11
    a = np.random.default_rng(8).integers(0, 256, n, dtype=np.uint8)
    t = torch.from_numpy(a).to("cuda:0")
    bad = N.SlotHandle.create_output("pushc/bad", 0, n)
    good = N.SlotHandle.create_output("pushc/good", 0, n)
    inp = N.SlotHandle.open_input("pushc/in", 1, n)
    N.set_copy_engine(N.ENGINE_PUSH)
    try:
        for s in (bad, good):
            s.pack(t.data_ptr(), n)
            s.commit()
        bad.fault(N.FAULT_CORRUPT_ONCE, (9 << 20) + 3)
        before = N.stats()["push_launches"]
        assert inp.transfer_chunk(bad, 0, 1 << 40) == n
        assert N.stats()["push_launches"] - before == 1
        with pytest.raises(N.ReadException):
            inp.finish(bad)
        assert inp.info().delivered == 0
        assert inp.transfer_chunk(good, 0, 1 << 40) == n
        inp.finish(good)
        assert tuple(inp.info().digest) == oracle.digest(a)
        assert np.array_equal(to_host(inp), a)
    finally:
        N.set_copy_engine(N.ENGINE_AUTO)
        for s in (bad, good, inp):
            s.close()


def test_both_directions_at_once_are_bit_exact(native_all):
    """the ring's situation: GPU 0 → GPU 1 and GPU 1 → GPU 0 in flight together on their own streams, each GPU pushing
    and being pushed into at the same time.  Under the auto policy the first transfer finds its sink GPU idle and is
    pulled; the second finds its sink GPU serving the first and is pushed."""
    n = (192 << 20) + 64
    ts, outs, inps, streams = [], [], [], []
    for g in (0, 1):
        t = torch.empty(n, dtype=torch.uint8, device=f"cuda:{g}")
        N.fill_pattern(g, t.data_ptr(), n, 31 + g)
        torch.cuda.synchronize(g)
        ts.append(t)
        o = N.SlotHandle.create_output(f"bidir/out{g}", g, n)
        o.pack(t.data_ptr(), n)
        o.commit()
        outs.append(o)
        streams.append(torch.cuda.Stream(device=g))
    try:
        for g in (0, 1):
            inps.append(N.SlotHandle.open_input(f"bidir/in{g}", g, n))
        for g in (0, 1):
            torch.cuda.synchronize(g)
        before = N.stats()
        for g in (0, 1):  # queue both before waiting for either
            assert inps[g].transfer_chunk(outs[1 - g], 0, n, int(streams[g].cuda_stream)) == n
        for g in (0, 1):
            inps[g].finish(outs[1 - g], int(streams[g].cuda_stream))
            assert tuple(inps[g].info().digest) == tuple(outs[1 - g].info().digest) == oracle.digest(oracle.pattern(n, 31 + 1 - g))
            from lzy_b200.devmem import view_bytes
            assert torch.equal(view_bytes(inps[g].device_ptr(), n, g).to(f"cuda:{1 - g}"), ts[1 - g])
        after = N.stats()
        assert after["push_launches"] - before["push_launches"] == 1 and after["dma_copies"] == before["dma_copies"]
        assert after["bytes_pulled_peer"] - before["bytes_pulled_peer"] == 2 * n
    finally:
        for s in outs + inps:
            s.close()


@pytest.mark.parametrize("chunk", [16 << 20, 0], ids=["host-driven-16MiB", "device-chained"])
def test_broadcast_chain_to_every_gpu(native_all, chunk):
    """C3 at test size: one producer on GPU 0, one consumer per GPU (GPU 0 included).  chunk = 0 is the
    device-chained kernel (one launch per GPU, progress counters in peer memory)."""
    n = (96 << 20) + 33
    t = torch.empty(n + 16, dtype=torch.uint8, device="cuda:0")
    N.fill_pattern(0, t.data_ptr(), n, 7)
    torch.cuda.synchronize(0)
    expect = oracle.pattern(n, 7)
    src = N.SlotHandle.create_output("bc/src", 0, n)
    dsts = [N.SlotHandle.open_input(f"bc/dst{g}", g, n) for g in range(NGPU)]
    try:
        src.pack(t.data_ptr(), n)
        dg = src.commit()
        assert dg == oracle.digest(expect)
        N.broadcast(src, dsts, chunk)
        for d in dsts:
            assert d.info().state == N.ST_READY and tuple(d.info().digest) == dg
            assert np.array_equal(to_host(d), expect)
    finally:
        src.close()
        for d in dsts:
            d.close()


@pytest.fixture(scope="module")
def lzy():
    rt = B200Runtime(placement="round_robin")
    yield Lzy(runtime=rt)
    rt.close()


def test_chain_round_robin_over_gpus(lzy):
    """C2's placement rule: op k on GPU k mod G; every edge crosses NVLink."""
    seen = []

    @op
    def hop(t: torch.Tensor) -> torch.Tensor:
        seen.append(t.device.index)
        return t

    x = torch.empty(4096, 4096, device="cuda:0").normal_()
    keep = x.clone()
    before = N.stats()["bytes_pulled_peer"]
    with lzy.workflow(f"wf{uuid.uuid4().hex}"):
        y = x
        for _ in range(8):
            y = hop(y)
    assert torch.equal(y.to("cuda:0"), keep)
    assert len(set(seen)) == min(NGPU, 8) and N.stats()["bytes_pulled_peer"] - before >= 6 * x.numel() * 4


def test_long_edges_through_the_runtime_ride_sm_kernels_only(lzy):
    """a 256 MiB edge between ops on different GPUs is carried by the hand-written kernels inside the slots runtime's
    download loop (not only when lzyb_transfer_chunk is called by hand) — pull or push as the auto policy decides, never the
    copy engine — and stays bit-exact"""
    @op
    def hop(t: torch.Tensor) -> torch.Tensor:
        return t

    x = torch.empty(8192, 8192, device="cuda:0").normal_()  # 256 MiB
    keep = x.clone()
    before = N.stats()
    with lzy.workflow(f"wf{uuid.uuid4().hex}"):
        y = hop(hop(x))
    assert torch.equal(y.to("cuda:0"), keep)
    after = N.stats()
    assert after["bytes_pulled_peer"] - before["bytes_pulled_peer"] >= x.numel() * 4  # at least one cross-GPU edge
    assert after["dma_copies"] == before["dma_copies"]


def test_diamond_mixed_payloads_across_gpus(lzy):
    """C5 at test size: source → branches (2 ops each) → sink; each edge carries a tensor and a dict."""
    @op
    def source() -> (torch.Tensor, Dict[str, int]):
        return torch.arange(1 << 20, dtype=torch.float32, device="cuda"), {"hops": 0}

    @op
    def branch(t: torch.Tensor, meta: Dict[str, int], k: int) -> (torch.Tensor, Dict[str, int]):
        return t * k, {"hops": meta["hops"] + 1, "k": k}

    @op
    def sink(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, ma: Dict[str, int], mb: Dict[str, int],
             mc: Dict[str, int]) -> (float, int):
        dev = a.device
        return float((a + b.to(dev) + c.to(dev)).sum().item()), ma["hops"] + mb["hops"] + mc["hops"]

    with lzy.workflow(f"wf{uuid.uuid4().hex}"):
        t, m = source()
        outs = []
        for k in (1, 2, 3):
            t1, m1 = branch(t, m, k)
            outs.append(branch(t1, m1, 1))
        total, hops = sink(outs[0][0], outs[1][0], outs[2][0], outs[0][1], outs[1][1], outs[2][1])
    n = 1 << 20
    assert total == float(6 * (n * (n - 1) // 2)) or abs(total - 6 * (n * (n - 1) / 2)) / (6 * n * n / 2) < 1e-6
    assert hops == 6


def test_locality_placement_follows_the_data():
    rt = B200Runtime(placement="locality")
    try:
        lz = Lzy(runtime=rt)
        where = []

        @op
        def here(t: torch.Tensor) -> int:
            where.append(t.device.index)
            return t.device.index

        with lz.workflow(f"wf{uuid.uuid4().hex}"):
            r = here(torch.ones(1 << 20, device="cuda:1"))
        assert r == 1 and where == [1]
    finally:
        rt.close()


@pytest.mark.skipif(NGPU < 4, reason="needs at least 4 GPUs (a producer and 3 consumers elsewhere)")
def test_fanout_to_many_gpus_uses_the_pull_chain():
    """C3 through the @op API: one output consumed on ≥ 3 other GPUs is delivered by one pull chain."""
    rt = B200Runtime(placement="round_robin", fanout_threshold=3)
    try:
        lz = Lzy(runtime=rt)

        @op
        def produce() -> torch.Tensor:
            return torch.arange(1 << 22, dtype=torch.int64, device="cuda")

        @op
        def consume(t: torch.Tensor, k: int) -> int:
            t.mul_(k)  # every consumer owns its copy
            return int(t[-1].item())

        with lz.workflow(f"wf{uuid.uuid4().hex}"):
            src = produce()
            outs = [consume(src, k) for k in range(1, NGPU)]
        assert [int(o) for o in outs] == [((1 << 22) - 1) * k for k in range(1, NGPU)]
        assert rt.stats.get("fanouts", 0) == 1
    finally:
        rt.close()
