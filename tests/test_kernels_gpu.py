"""GPU parity tests proper: the CUDA data plane (through the C ABI) against the CPU oracle.

Shapes follow the reference's own slot tests (lzy/slots/src/test/java/ai/lzy/slots/SlotsTest.java):
"Hello" (:130-187), 200 MiB of random bytes compared byte for byte (testLargeData, :628-693),
2 MiB with a failure after 1 MiB and an offset resume (testRestoreReadFromOffset, :803-864).
Everything here is integer/byte work, so the bar is bit-exact.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from lzy_b200.devmem import view_bytes  # noqa: E402

SIZES = [0, 1, 7, 8, 15, 16, 17, 255, 4096, 4097, 65535, (2 << 20) - 1, 2 << 20, (2 << 20) + 1, (8 << 20) + 13]


def dev_bytes(a: np.ndarray, pad_front: int = 0):
    """numpy u8 → cuda tensor; pad_front shifts the data pointer to test misaligned sources."""
    t = torch.empty(a.size + pad_front + 64, dtype=torch.uint8, device="cuda:0")
    v = t[pad_front:pad_front + a.size]
    if a.size:
        v.copy_(torch.from_numpy(a))
    torch.cuda.synchronize()
    return t, v


def slot_bytes(n, slot) -> np.ndarray:
    out = np.empty(slot.nbytes, dtype=np.uint8)
    if slot.nbytes:
        slot.copy_to_host(out.ctypes.data, slot.nbytes)
    return out


def rnd(n, seed=0):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)


@pytest.mark.parametrize("nbytes", SIZES)
def test_digest_device_matches_oracle(native, nbytes):
    a = rnd(nbytes, nbytes)
    _, v = dev_bytes(a)
    assert native.digest_device(0, v.data_ptr() if nbytes else 0, nbytes) == oracle.digest(a)


@pytest.mark.parametrize("shift", [0, 1, 2, 4, 8, 12])
@pytest.mark.parametrize("nbytes", [1, 17, 4099, (1 << 20) + 5])
def test_pack_any_alignment(native, nbytes, shift):
    a = rnd(nbytes, nbytes + shift)
    _, v = dev_bytes(a, pad_front=shift)
    s = native.SlotHandle.create_output(f"pack-{nbytes}-{shift}", 0, nbytes)
    try:
        s.pack(v.data_ptr(), nbytes)
        assert s.commit() == oracle.digest(a)
        assert np.array_equal(slot_bytes(native, s), a)
        hd = s.header()
        assert bytes(hd.magic) == b"LZYB200\0" and hd.nbytes == nbytes and hd.flags & 1
        assert (int(hd.digest[0]), int(hd.digest[1])) == oracle.digest(a)
    finally:
        s.close()


def test_commit_in_place_scan(native):
    """Producer writes straight into the slot (the fifo analogue) and commit() scans it."""
    n = (3 << 20) + 7
    a = rnd(n, 5)
    s = native.SlotHandle.create_output("inplace", 0, n)
    try:
        from lzy_b200.devmem import view_bytes
        view_bytes(s.device_ptr(), n, 0).copy_(torch.from_numpy(a).cuda())
        torch.cuda.synchronize()
        assert s.commit() == oracle.digest(a)
        assert np.array_equal(slot_bytes(native, s), a)
    finally:
        s.close()


def test_hello_slot_to_slot(native):
    """SlotsTest.testSimple's payload."""
    data = np.frombuffer(b"Hello", dtype=np.uint8)
    _, v = dev_bytes(data)
    out = native.SlotHandle.create_output("hello-out", 0, 5)
    inp = native.SlotHandle.open_input("hello-in", 0, 5)
    try:
        out.pack(v.data_ptr(), 5)
        out.commit()
        assert inp.transfer_chunk(out, 0, 1 << 30) == 5
        assert inp.transfer_chunk(out, 5, 1 << 30) == native.EOF
        inp.finish(out)
        assert slot_bytes(native, inp).tobytes() == b"Hello"
        assert inp.info().state == native.ST_READY
        # re-publication as "<slot>-out" (InputSlot.java:160-168): same frame, now readable
        inp.to_output()
        assert inp.read(0, 2 << 20) == b"Hello" and inp.read(5, 2 << 20) is None
    finally:
        out.close()
        inp.close()


@pytest.mark.parametrize("engine", ["ldg", "tma"])
@pytest.mark.parametrize("nbytes", [(2 << 20) * 3 + 4096 + 5, 64 << 20])
def test_pull_chunked_2mib_matches_stream_oracle(native, nbytes, engine):
    """Consumer pulls in 2 MiB chunks exactly like Read → transferChunkTo; delivered bytes and
    chunk count equal the oracle's restatement of OutputSlot.ReadThread."""
    native.set_copy_engine(native.ENGINE_TMA if engine == "tma" else native.ENGINE_LDG)
    try:
        a = rnd(nbytes, 11)
        _, v = dev_bytes(a)
        out = native.SlotHandle.create_output("chunk-out", 0, nbytes)
        inp = native.SlotHandle.open_input("chunk-in", 0, nbytes)
        try:
            out.pack(v.data_ptr(), nbytes)
            dg = out.commit()
            ref_dst = np.zeros(nbytes, dtype=np.uint8)
            ref_chunks, ref_delivered = oracle.stream_copy(a, ref_dst)
            chunks, off = 0, 0
            while True:
                n = inp.transfer_chunk(out, off, 2 << 20)
                if n == native.EOF:
                    break
                off += n
                chunks += 1
            inp.finish(out)
            assert (chunks, off) == (ref_chunks, ref_delivered)
            assert np.array_equal(slot_bytes(native, inp), ref_dst)
            i = inp.info()
            assert (int(i.digest[0]), int(i.digest[1])) == dg == oracle.digest(a)
        finally:
            out.close()
            inp.close()
    finally:
        native.set_copy_engine(native.ENGINE_AUTO)


@pytest.mark.parametrize("engine", ["ldg", "tma"])
def test_large_data_200mib(native, engine):
    """SlotsTest.testLargeData: 200 × 1 MiB random bytes, Files.mismatch == -1."""
    native.set_copy_engine(native.ENGINE_TMA if engine == "tma" else native.ENGINE_LDG)
    try:
        n = 200 << 20
        a = rnd(n, 2024)
        _, v = dev_bytes(a)
        out = native.SlotHandle.create_output("large-out", 0, n)
        inp = native.SlotHandle.open_input("large-in", 0, n)
        try:
            out.pack(v.data_ptr(), n)
            out.commit()
            assert inp.transfer_chunk(out, 0, 1 << 40) == n
            inp.finish(out)
            got = slot_bytes(native, inp)
            assert np.array_equal(got, a)
            assert native.digest_device(0, inp.device_ptr(), n) == oracle.digest(a)
        finally:
            out.close()
            inp.close()
    finally:
        native.set_copy_engine(native.ENGINE_AUTO)


def test_restore_read_from_offset(native):
    """SlotsTest.testRestoreReadFromOffset: 2 MiB, producer fails after 1 MiB; the first MiB is
    already in place; a new transfer resumes at the delivered offset and the result is exact."""
    n = 2 << 20
    a = rnd(n, 77)
    _, v = dev_bytes(a)
    out = native.SlotHandle.create_output("restore-out", 0, n)
    inp = native.SlotHandle.open_input("restore-in", 0, n)
    try:
        out.pack(v.data_ptr(), n)
        out.commit()
        out.fault(native.FAULT_PEER_AFTER, 1 << 20)
        off = 0
        with pytest.raises(native.ReadException):
            while True:
                moved = inp.transfer_chunk(out, off, 512 << 10)
                assert moved != native.EOF
                off += moved
        assert off == 1 << 20 == inp.info().delivered
        ref = np.zeros(n, dtype=np.uint8)
        _, ref_delivered = oracle.stream_copy(a, ref, 0, fail_after=1 << 20)
        # (the reference's 2 MiB buffer would have read past the failure point in one read();
        #  the in-memory test backend fails *between* reads, so exactly 1 MiB lands)
        part = np.empty(1 << 20, dtype=np.uint8)
        inp.copy_to_host(part.ctypes.data, 1 << 20)
        assert np.array_equal(part, a[: 1 << 20])
        # TransferFailed → same peer again with the failure cleared, resume from offset
        out.fault(native.FAULT_NONE)
        while True:
            moved = inp.transfer_chunk(out, off, 512 << 10)
            if moved == native.EOF:
                break
            off += moved
        inp.finish(out)
        assert off == n
        assert np.array_equal(slot_bytes(native, inp), a)
    finally:
        out.close()
        inp.close()


def test_corrupt_peer_is_a_read_exception_then_refetch(native):
    n = (5 << 20) + 3
    a = rnd(n, 3)
    _, v = dev_bytes(a)
    bad = native.SlotHandle.create_output("bad-peer", 0, n)
    good = native.SlotHandle.create_output("good-peer", 0, n)
    inp = native.SlotHandle.open_input("victim", 0, n)
    try:
        for s in (bad, good):
            s.pack(v.data_ptr(), n)
            s.commit()
        bad.fault(native.FAULT_CORRUPT_ONCE, 12345)  # the peer now serves one wrong byte
        assert inp.transfer_chunk(bad, 0, 1 << 40) == n
        with pytest.raises(native.ReadException):
            inp.finish(bad)
        assert inp.info().delivered == 0 and inp.info().state == native.ST_EMPTY
        assert inp.transfer_chunk(good, 0, 1 << 40) == n
        inp.finish(good)
        assert np.array_equal(slot_bytes(native, inp), a)
    finally:
        bad.close()
        good.close()
        inp.close()


def test_legacy_read_stream_2mib_chunks_and_eos(native):
    """OutputSlot.ReadThread semantic for legacy gRPC consumers: bytes [offset, end) in ≤ 2 MiB
    chunks then EOS (OutputSlot.java:238-265)."""
    n = (5 << 20) + 100
    a = rnd(n, 9)
    _, v = dev_bytes(a)
    out = native.SlotHandle.create_output("legacy", 0, n)
    try:
        out.pack(v.data_ptr(), n)
        out.commit()
        for start in (0, 1 << 20, n - 1, n):
            got, off, sizes = bytearray(), start, []
            while True:
                chunk = out.read(off, 2 << 20)
                if chunk is None:
                    break
                got += chunk
                off += len(chunk)
                sizes.append(len(chunk))
            assert bytes(got) == a[start:].tobytes()
            assert all(s == 2 << 20 for s in sizes[:-1])
    finally:
        out.close()


def test_pack_2d_gather(native):
    rows, cols = 777, 1024 + 64  # fp32 matrix, take a column window → pitch != row_bytes
    m = np.random.default_rng(1).standard_normal((rows, cols)).astype(np.float32)
    t = torch.from_numpy(m).cuda()
    view = t[:, 16:16 + 1024]
    row_bytes, pitch = 1024 * 4, cols * 4
    expect = np.ascontiguousarray(m[:, 16:16 + 1024]).view(np.uint8).reshape(-1)
    flat = np.concatenate([m.view(np.uint8).reshape(-1), np.zeros(pitch, np.uint8)])
    assert np.array_equal(expect, oracle.gather2d(flat[64:], rows, row_bytes, pitch))
    s = native.SlotHandle.create_output("gather", 0, rows * row_bytes)
    try:
        s.set_meta(native.DT_F32, (rows, 1024))
        s.pack_2d(view.data_ptr(), rows, row_bytes, pitch)
        assert s.commit() == oracle.digest(expect)
        assert np.array_equal(slot_bytes(native, s), expect)
        hd = s.header()
        assert hd.dtype == native.DT_F32 and hd.ndim == 2 and list(hd.shape[:2]) == [rows, 1024]
    finally:
        s.close()


def test_fill_pattern_matches_oracle(native):
    for n in (16, 4096 + 8, (4 << 20) + 3):
        t = torch.empty(n + 16, dtype=torch.uint8, device="cuda:0")
        native.fill_pattern(0, t.data_ptr(), n, 7)
        torch.cuda.synchronize()
        assert np.array_equal(t[:n].cpu().numpy(), oracle.pattern(n, 7))


def test_host_ring_frames_pickles(native):
    import pickle
    objs = [b"", b"x", pickle.dumps({"a": 1, "b": [1, 2, 3], "s": "kek" * 1000}), bytes(rnd(70000, 4))]
    offs = [native.frame_host_object(o) for o in objs]
    for o, off in zip(objs, offs):
        hd = native.ring_frame_info(off)
        assert hd.nbytes == len(o) and hd.dtype == native.DT_PICKLE
        assert (int(hd.digest[0]), int(hd.digest[1])) == oracle.digest(o)
        assert native.ring_read(off) == o
    for off in offs:
        native.ring_release(off)


def test_pack_host_and_input_write_host(native):
    n = (3 << 20) + 33
    a = rnd(n, 21)
    s = native.SlotHandle.create_output("from-host", 0, n)
    inp = native.SlotHandle.open_input("host-fed", 0, n)
    try:
        s.pack_host(a.ctypes.data, n)
        dg = s.commit()
        assert dg == oracle.digest(a) and np.array_equal(slot_bytes(native, s), a)
        off = 0
        while off < n:
            step = min(1 << 20, n - off)
            assert inp.write_host(off, a.ctypes.data + off, step) == step
            off += step
        inp.finish_expect(dg)
        assert np.array_equal(slot_bytes(native, inp), a)
    finally:
        s.close()
        inp.close()


def test_digest_is_chunk_composable(native):
    """Size-independent property used at BASELINE sizes: partial digests over disjoint ranges add up."""
    n = 32 << 20
    a = rnd(n, 8)
    acc = (0, 0)
    for off in range(0, n, 4 << 20):
        acc = oracle.digest_partial(a[off:off + (4 << 20)], off, acc)
    assert oracle.digest_finalize(acc[0], acc[1], n) == oracle.digest(a)
    _, v = dev_bytes(a)
    assert native.digest_device(0, v.data_ptr(), n) == oracle.digest(a)


def test_stats_count_our_kernels(native):
    before = native.stats()["kernel_launches"]
    t = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda:0")
    native.digest_device(0, t.data_ptr(), 1 << 20)
    assert native.stats()["kernel_launches"] == before + 1


def test_sole_consumer_on_the_same_gpu_gets_the_frame_not_a_copy(native):
    """`FileInputBackend.toOutput()` hands a file over by ownership (FileInputBackend.java:38-45); the
    same-GPU producer → consumer edge does the same once the producer's owner has declared a sole consumer and nobody
    else holds the frame.  No kernel runs; the consumer owns the committed frame, the producer reads MOVED."""
    N = native
    n = (6 << 20) + 5
    a = np.random.default_rng(11).integers(0, 256, n, dtype=np.uint8)
    t = torch.from_numpy(a).cuda()
    out = N.SlotHandle.create_output("mv/out", 0, n)
    out.set_meta(N.DT_U8, (n,))
    out.pack(t.data_ptr(), n)
    dg = out.commit()
    src_ptr = out.device_ptr()
    inp = N.SlotHandle.open_input("mv/in", 0, n)
    try:
        out.set_movable(True)
        before = N.stats()
        assert inp.transfer_chunk(out, 0, 1 << 40) == n
        inp.finish(out)
        after = N.stats()
        assert after["kernel_launches"] == before["kernel_launches"]  # nothing was copied, nothing re-scanned
        assert after["frames_moved"] - before["frames_moved"] == 1
        assert inp.device_ptr() == src_ptr and out.device_ptr() != src_ptr
        assert inp.info().state == N.ST_READY and tuple(inp.info().digest) == dg == oracle.digest(a)
        assert inp.info().dtype == N.DT_U8 and inp.info().ndim == 1 and inp.info().shape[0] == n
        assert out.info().state == N.ST_MOVED
        assert np.array_equal(view_bytes(inp.device_ptr(), n, 0).cpu().numpy(), a)
        hdr = inp.header()
        assert bytes(hdr.magic)[:7] == b"LZYB200" and int(hdr.nbytes) == n and tuple(hdr.digest) == dg
        assert inp.transfer_chunk(out, n, 1 << 20) == N.EOF
        late = N.SlotHandle.open_input("mv/late", 0, n)
        with pytest.raises(N.ReadException):  # a second consumer that was promised not to exist
            late.transfer_chunk(out, 0, 1 << 40)
        late.close()
        # the consumer can re-publish what it now owns (InputSlot "-out" re-publication) and serve it by copy
        inp.to_output()
        again = N.SlotHandle.open_input("mv/again", 0, n)
        assert again.transfer_chunk(inp, 0, 1 << 40) == n
        again.finish(inp)
        assert np.array_equal(view_bytes(again.device_ptr(), n, 0).cpu().numpy(), a)
        again.close()
    finally:
        out.close()
        inp.close()


@pytest.mark.parametrize("why", ["not declared", "second holder", "partial request"])
def test_the_copy_stays_whenever_someone_else_may_read_the_frame(native, why):
    """snapshot-at-call-time (pylzy tests/api/v1/test_entries.py:331-384): a frame that a proxy, the frame store or a
    second consumer can still ask for is never handed over"""
    N = native
    n = (3 << 20) + 32
    a = np.random.default_rng(12).integers(0, 256, n, dtype=np.uint8)
    t = torch.from_numpy(a).cuda()
    out = N.SlotHandle.create_output(f"mvc/{why}/out", 0, n)
    out.pack(t.data_ptr(), n)
    out.commit()
    inp = N.SlotHandle.open_input(f"mvc/{why}/in", 0, n)
    try:
        if why != "not declared":
            out.set_movable(True)
        if why == "second holder":
            out.retain()
        src_ptr = out.device_ptr()
        before = N.stats()["frames_moved"]
        step = (1 << 20) if why == "partial request" else (1 << 40)
        off = 0
        while True:
            moved = inp.transfer_chunk(out, off, step)
            if moved == N.EOF:
                break
            off += moved
        inp.finish(out)
        assert N.stats()["frames_moved"] == before
        assert out.info().state == N.ST_READY and out.device_ptr() == src_ptr and inp.device_ptr() != src_ptr
        assert np.array_equal(view_bytes(inp.device_ptr(), n, 0).cpu().numpy(), a)
        assert np.array_equal(view_bytes(out.device_ptr(), n, 0).cpu().numpy(), a)
    finally:
        if why == "second holder":
            out.close()
        out.close()
        inp.close()


def test_slot_slab_exhaustion_is_an_error_not_a_hang(native):
    """with every per-device slot control entry taken, slot creation used to free the frame while
    holding the device mutex that free_frame takes — a self-deadlock.  Provoked through the reserve hook."""
    N = native
    N.lib.lzyb_fault_inject(0, N.FAULT_SLAB_RESERVE, 1 << 20)
    try:
        live = N.stats()["slots_live"]
        with pytest.raises(N.LzybError, match="too many live slots"):
            N.SlotHandle.create_output("slab/out", 0, 4 << 20)
        assert N.stats()["slots_live"] == live
    finally:
        N.lib.lzyb_fault_inject(0, N.FAULT_SLAB_RESERVE, 0)
    s = N.SlotHandle.create_output("slab/ok", 0, 4 << 20)  # and the frame it had taken went back to the pool
    s.close()


@pytest.mark.parametrize("store", [True, False], ids=["copy", "scan"])
def test_tma_ring_recycling_under_stress(native, store):
    """compute-sanitizer's racecheck flags write-after-read hazards on the TMA ring's shared-memory
    stages that the full/empty mbarrier pair is there to order (profiles/r01_sanitizer_racecheck_hazards.txt).  The
    argument is checked here the hard way: the shallowest ring the kernel supports (2 stages of 8 KiB) on 4 CTAs, so every
    CTA recycles each stage thousands of times per transfer, hundreds of transfers of awkward sizes — a stage released one
    phase early or loaded one phase late would deliver a wrong byte or a wrong digest."""
    N = native
    N.set_copy_engine(N.ENGINE_TMA)
    N.set_tma_config(6)
    N.lib.lzyb_set_debug_flags(2)
    try:
        for i, n in enumerate([(1 << 20), (1 << 20) + 8192 * 3 + 48, (3 << 20) + 16, (5 << 20) + 8192 + 1, 16 << 20]):
            a = rnd(n, seed=100 + i)
            t, v = dev_bytes(a)
            want = oracle.digest(a)
            for rep in range(20):
                if store:
                    out = N.SlotHandle.create_output(f"stress/out{i}/{rep}", 0, n)
                    out.pack(v.data_ptr(), n)
                    out.commit()
                    inp = N.SlotHandle.open_input(f"stress/in{i}/{rep}", 0, n)
                    assert inp.transfer_chunk(out, 0, 1 << 40) == n  # the TMA ring, copying
                    inp.finish(out)
                    assert tuple(inp.info().digest) == want
                    if rep % 5 == 0:
                        assert np.array_equal(slot_bytes(n, inp), a)
                    inp.close()
                    out.close()
                else:
                    assert N.digest_device(0, v.data_ptr(), n) == want  # the TMA ring, scanning
    finally:
        N.lib.lzyb_set_debug_flags(0)
        N.set_tma_config(1)
        N.set_copy_engine(N.ENGINE_AUTO)
