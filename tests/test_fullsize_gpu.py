"""BASELINE.json's full sizes, checked through size-independent properties (the oracle would take
minutes at these sizes): digest-of-chunks composition, encode→transfer→decode round trips compared on
the device, idempotence of re-publication, and the counter-based pattern whose any 4 KiB window the
oracle can regenerate on its own (word i depends on (seed, i) only, so
pattern(seed + k)[j] == pattern(seed)[j + 8k])."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

GIB = 1 << 30


def window_check(slot, nbytes, seed, rng):
    """random 4 KiB windows of the payload equal the oracle's independently generated pattern"""
    for off in [0, nbytes - 4096] + [int(rng.integers(0, nbytes // 16 - 256)) * 16 for _ in range(6)]:
        got = np.empty(4096, dtype=np.uint8)
        slot.copy_to_host(got.ctypes.data, 4096, off)
        assert np.array_equal(got, oracle.pattern(4096, seed + off // 8)), f"window at {off} differs"


def test_one_gib_chain_of_eight_hops_is_bit_exact(native):
    """configs[1] at full size: 8 hops of 1 GiB; every hop's digest equals the first; the final payload
    equals the source on the device."""
    n = GIB
    src = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    native.fill_pattern(0, src.data_ptr(), n, 1234)
    torch.cuda.synchronize()
    first = native.SlotHandle.create_output("full/ingest", 0, n)
    first.pack(src.data_ptr(), n)
    dg = first.commit()
    live, prev = [first], first
    for k in range(8):
        inp = native.SlotHandle.open_input(f"full/in{k}", 0, n)
        assert inp.transfer_chunk(prev, 0, n) == n
        inp.finish(prev)
        out = native.SlotHandle.create_output(f"full/out{k}", 0, n)
        out.pack(inp.device_ptr(), n)
        assert out.commit() == dg
        live += [inp, out]
        prev = out
    from lzy_b200.devmem import view_bytes
    assert torch.equal(view_bytes(prev.device_ptr(), n, 0), src)
    assert native.digest_device(0, prev.device_ptr(), n) == dg
    window_check(prev, n, 1234, np.random.default_rng(0))
    for s in live:
        s.close()


def test_resume_at_full_size_with_mid_transfer_peer_failure(native):
    n = GIB
    src = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    native.fill_pattern(0, src.data_ptr(), n, 99)
    a = native.SlotHandle.create_output("full/a", 0, n)
    b = native.SlotHandle.create_output("full/b", 0, n)
    inp = native.SlotHandle.open_input("full/c", 0, n)
    try:
        for s in (a, b):
            s.pack(src.data_ptr(), n)
        assert a.commit() == b.commit()
        a.fault(native.FAULT_PEER_AFTER, 640 << 20)
        off = 0
        with pytest.raises(native.ReadException):
            while True:
                off += inp.transfer_chunk(a, off, 256 << 20)
        assert off == 640 << 20
        while True:  # the channel manager hands out another holder of the same data; resume at `off`
            moved = inp.transfer_chunk(b, off, 256 << 20)
            if moved == native.EOF:
                break
            off += moved
        inp.finish(b)
        assert tuple(inp.info().digest) == a.commit()
        window_check(inp, n, 99, np.random.default_rng(1))
    finally:
        for s in (a, b, inp):
            s.close()


def test_four_gib_pattern_slot_commit_and_republish(native):
    """C3's payload: a 4 GiB pattern written in place, committed by a scan, pulled once, re-published
    (toOutput) and read back through the legacy path at a few offsets."""
    n = 4 * GIB
    out = native.SlotHandle.create_output("full/4g", 0, n)
    inp = native.SlotHandle.open_input("full/4g-in", 0, n)
    try:
        native.fill_pattern(0, out.device_ptr(), n, 7)
        dg = out.commit()
        assert inp.transfer_chunk(out, 0, n) == n
        inp.finish(out)
        assert tuple(inp.info().digest) == dg
        inp.to_output()
        for off in (0, 3 * GIB + 12344, n - 96):
            got = inp.read(off, 4096)
            assert len(got) == min(4096, n - off)
            assert got == oracle.pattern(4096, 7 + off // 8)[: len(got)].tobytes()
        assert inp.read(n, 4096) is None
    finally:
        out.close()
        inp.close()
