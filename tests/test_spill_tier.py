"""The spill tier below the HBM frame store (include/lzy_b200_host.h: lzyb_frame_store_set_spill / spill / load)
— SURVEY §8(f) row 3: a frame that left HBM stays fetchable as a BACKUP storage producer
(channel-manager PeerDao.java:44-53), cached results survive under stable URIs (core/workflow.py:247-281).
Host logic only, over the in-memory slot backend; the HBM side of the same calls is tests/test_spill_gpu.py."""
import os
import struct

import pytest

from lzy_b200 import _host as H
from lzy_b200 import _native as N
from lzy_b200._host import DIR_INPUT, DIR_OUTPUT, ChannelManager, Slots
from tests.support import hostmem


@pytest.fixture
def env(tmp_path):
    lib = hostmem.install()
    cm = ChannelManager(seed=4)
    H.pump_start(cm)
    slots = Slots(cm, 0, "lzyb://worker-0")
    H.frame_store_set_spill(str(tmp_path / "spill"), 0)
    yield cm, slots, lib, tmp_path / "spill"
    H.frame_store_clear("")
    H.frame_store_set_spill(None)
    slots.close()
    H.pump_stop(cm)
    cm.close()
    hostmem.uninstall()
    assert lib.hostmem_live_slots() == 0, "a slot frame leaked"


def produce_to_store(cm, slots, lib, uri, data, dtype=0, shape=()):
    """an op output whose channel's storage consumer is `uri`: after afterExecution the frame store holds it"""
    chan = cm.get_or_create("u", "wf", "exec", consumer_uri=uri)
    ctx = slots.context("exec", "p-" + uri[-6:])
    ctx.add_slot("out", chan, DIR_OUTPUT)
    ctx.before_execution()
    h = ctx.output_open("out", len(data))
    if shape:
        hostmem.set_meta(h, dtype, shape)
    hostmem.write(h, data)
    ctx.output_done("out")
    ctx.after_execution(5000)
    return ctx, chan


def test_spill_file_is_header_plus_payload_and_loads_back_verified(env):
    cm, slots, lib, d = env
    data = os.urandom((3 << 20) + 7)
    ctx, _ = produce_to_store(cm, slots, lib, "hbm://t/a", data, dtype=N.DT_F32, shape=(3, 5, 7))
    assert not H.frame_store_spilled("hbm://t/a")
    H.frame_store_spill("hbm://t/a")
    assert H.frame_store_spilled("hbm://t/a")
    (f,) = list(d.iterdir())
    raw = f.read_bytes()
    assert len(raw) == 128 + len(data) and raw[:8] == b"LZYB200\0" and raw[128:] == data
    version, dtype, nbytes, ndim, flags = struct.unpack_from("<IIQII", raw, 8)
    assert (version, dtype, nbytes, ndim, flags & 1) == (N.lib.lzyb_abi_version(), N.DT_F32, len(data), 3, 1)
    assert struct.unpack_from("<3Q", raw, 32) == (3, 5, 7)
    resident = H.frame_store_get("hbm://t/a")
    digest = tuple(hostmem.info(resident).digest)
    assert struct.unpack_from("<2Q", raw, 96) == digest

    ctx.close()
    H.frame_store_drop("hbm://t/a")  # leaves HBM …
    assert lib.hostmem_live_slots() == 0
    h = H.frame_store_get("hbm://t/a")  # … and comes back from the tier, once, verified against the header
    assert h is not None and h == H.frame_store_get("hbm://t/a")
    info = hostmem.info(h)
    assert hostmem.peek(h, len(data)) == data and tuple(info.digest) == digest
    assert (info.dtype, info.ndim, tuple(info.shape[:3])) == (N.DT_F32, 3, (3, 5, 7))


def test_evicted_frame_is_still_a_storage_producer_for_later_consumers(env):
    """the reference's BACKUP semantics: the op's worker is gone, the channel's storage peer serves the data"""
    cm, slots, lib, d = env
    data = os.urandom(1 << 20)
    ctx, chan = produce_to_store(cm, slots, lib, "hbm://t/evicted", data)
    H.frame_store_spill("hbm://t/evicted")
    ctx.close()                              # the slot producer unbinds
    H.frame_store_drop("hbm://t/evicted")    # and the frame leaves "HBM"
    cons = slots.context("exec", "late")
    cons.add_slot("in", chan, DIR_INPUT)
    cons.before_execution(10000)
    assert hostmem.peek(cons.input_handle("in"), len(data)) == data
    cons.close()


def test_corrupt_spill_file_is_a_read_error_not_wrong_data(env):
    cm, slots, lib, d = env
    data = os.urandom(300_000)
    ctx, chan = produce_to_store(cm, slots, lib, "hbm://t/rot", data)
    H.frame_store_spill("hbm://t/rot")
    ctx.close()
    H.frame_store_drop("hbm://t/rot")
    (f,) = list(d.iterdir())
    raw = bytearray(f.read_bytes())
    raw[128 + 1234] ^= 0x40
    f.write_bytes(bytes(raw))
    with pytest.raises(N.LzybError, match="unusable"):
        H.frame_store_get("hbm://t/rot")
    assert lib.hostmem_live_slots() == 0  # the half-filled frame was returned


def test_foreign_file_without_header_is_a_bytes_payload(env):
    """an object some other tool put under the URI's name (no frame magic): payload = the file"""
    cm, slots, lib, d = env
    d.mkdir(exist_ok=True)
    name = "hbm%3a%2f%2ft%2fforeign.lzyb"
    (d / name).write_bytes(b"plain old bytes")
    assert H.frame_store_spilled("hbm://t/foreign")
    h = H.frame_store_get("hbm://t/foreign")
    assert hostmem.peek(h, 64) == b"plain old bytes"
    H.frame_store_unspill("hbm://t/foreign")
    assert not H.frame_store_spilled("hbm://t/foreign") and not (d / name).exists()


def test_long_uris_get_bounded_distinct_file_names(env):
    cm, slots, lib, d = env
    base = "hbm://bucket/lzy_runs/wf/ops/fn/0.0.1/" + "_".join(["d41d8cd98f00b204e9800998ecf8427e"] * 12)
    ctxs = []
    for i in range(2):
        ctx, _ = produce_to_store(cm, slots, lib, f"{base}/return_{i}", bytes([i]) * 100)
        H.frame_store_spill(f"{base}/return_{i}")
        ctxs.append(ctx)
    names = sorted(p.name for p in d.iterdir())
    assert len(names) == 2 and all(len(n) < 255 for n in names)
    for ctx in ctxs:
        ctx.close()
    for i in range(2):
        H.frame_store_drop(f"{base}/return_{i}")
        assert hostmem.peek(H.frame_store_get(f"{base}/return_{i}"), 100) == bytes([i]) * 100


def test_without_a_spill_directory_the_tier_is_off(env):
    cm, slots, lib, d = env
    H.frame_store_set_spill(None)
    ctx, _ = produce_to_store(cm, slots, lib, "hbm://t/x", b"abc" * 100)
    with pytest.raises(N.LzybError):
        H.frame_store_spill("hbm://t/x")
    assert not H.frame_store_spilled("hbm://t/x")
    ctx.close()
    H.frame_store_drop("hbm://t/x")
    assert H.frame_store_get("hbm://t/x") is None


def test_concurrent_consumers_of_one_spilled_uri_share_a_single_load(env):
    import threading
    cm, slots, lib, d = env
    data = os.urandom(2 << 20)
    ctx, chan = produce_to_store(cm, slots, lib, "hbm://t/shared", data)
    H.frame_store_spill("hbm://t/shared")
    ctx.close()
    H.frame_store_drop("hbm://t/shared")
    assert lib.hostmem_live_slots() == 0
    got, errors = [], []

    def fetch():
        try:
            got.append(H.frame_store_get("hbm://t/shared"))
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=fetch) for _ in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors and len(set(got)) == 1 and got[0] is not None
    assert lib.hostmem_live_slots() == 1  # one frame, not eight

    # … and several input slots bound to the channel at once are all served from it
    ctxs = []
    for i in range(4):
        c = slots.context("exec", f"c{i}")
        c.add_slot(f"in{i}", chan, DIR_INPUT)
        c.prefetch()
        ctxs.append(c)
    for i, c in enumerate(ctxs):
        c.before_execution(10000)
        assert hostmem.peek(c.input_handle(f"in{i}"), len(data)) == data
    for c in ctxs:
        c.close()


# ---- "file://" storage peers: the durable, box-wide storage --------------------------------------------------------
def test_file_storage_peer_upload_download_meta_and_corruption(env, tmp_path):
    """an op output whose channel's storage consumer is a file:// URI lands as [header][payload]; any later consumer
    (any process of the box) reads it back with dtype / shape / digest; a rotten file is a read failure"""
    cm, slots, lib, d = env
    data = os.urandom((1 << 20) + 3)
    uri = f"file://{tmp_path}/store/deep/obj.lzyb"
    chan = cm.get_or_create("u", "wf", "exec-f", consumer_uri=uri)
    out = slots.context("exec-f", "p")
    out.add_slot("/f/out", chan, DIR_OUTPUT)
    out.before_execution()
    h = out.output_open("/f/out", len(data))
    hostmem.set_meta(h, N.DT_I64, (7, 11))
    hostmem.write(h, data)
    out.output_done("/f/out")
    out.after_execution(10000)
    raw = (tmp_path / "store/deep/obj.lzyb").read_bytes()
    assert raw[:8] == b"LZYB200\0" and raw[128:] == data
    out.close()  # the slot producer is gone: the file is all there is

    cons = slots.context("exec-f", "c")
    cons.add_slot("/f/in", chan, DIR_INPUT)
    cons.before_execution(10000)
    got = cons.input_handle("/f/in")
    info = hostmem.info(got)
    assert hostmem.peek(got, len(data)) == data and (info.dtype, info.ndim, tuple(info.shape[:2])) == (N.DT_I64, 2, (7, 11))
    cons.close()

    rotten = bytearray(raw)
    rotten[128 + 99] ^= 1
    (tmp_path / "store/deep/obj.lzyb").write_bytes(bytes(rotten))
    bad = slots.context("exec-f", "c2")
    bad.add_slot("/f/in2", chan, DIR_INPUT)
    with pytest.raises(N.LzybError, match="No more producers|Cannot get data"):
        bad.before_execution(10000)  # the only producer keeps failing its digest: the channel is given up
    bad.close()


def test_pre_existing_plain_file_is_a_storage_producer(env, tmp_path):
    cm, slots, lib, d = env
    p = tmp_path / "input.bin"
    p.write_bytes(b"bytes somebody else wrote")
    chan = cm.get_or_create("u", "wf", "exec-f2", producer_uri=f"file://{p}")
    cons = slots.context("exec-f2", "c")
    cons.add_slot("/f2/in", chan, DIR_INPUT)
    cons.before_execution(10000)
    assert hostmem.peek(cons.input_handle("/f2/in"), 64) == b"bytes somebody else wrote"
    cons.close()
    missing = cm.get_or_create("u", "wf", "exec-f3", producer_uri=f"file://{tmp_path}/nope")
    lost = slots.context("exec-f3", "c")
    lost.add_slot("/f3/in", missing, DIR_INPUT)
    with pytest.raises(N.LzybError):
        lost.before_execution(10000)
    lost.close()


@pytest.mark.parametrize("damage", ["truncated", "nbytes lies", "impossible rank"])
def test_a_frame_file_that_contradicts_itself_is_rotten_not_foreign(env, damage):
    """a file that carries the frame magic but whose header does not fit the file (truncated payload, wrong
    nbytes, ndim > 8) used to be served as a *foreign* object — header + partial payload as DT_BYTES under a freshly
    computed digest.  It is a read error, like any other rotten file."""
    cm, slots, lib, d = env
    data = os.urandom(200_000)
    ctx, chan = produce_to_store(cm, slots, lib, "hbm://t/selfcontra", data, dtype=N.DT_U8, shape=(200_000,))
    H.frame_store_spill("hbm://t/selfcontra")
    ctx.close()
    H.frame_store_drop("hbm://t/selfcontra")
    (f,) = list(d.iterdir())
    raw = bytearray(f.read_bytes())
    if damage == "truncated":
        raw = raw[: 128 + 150_000]
    elif damage == "nbytes lies":
        struct.pack_into("<Q", raw, 16, len(data) - 16)
    else:
        struct.pack_into("<I", raw, 24, 9)
    f.write_bytes(bytes(raw))
    with pytest.raises(N.LzybError, match="corrupt frame file"):
        H.frame_store_get("hbm://t/selfcontra")
    assert lib.hostmem_live_slots() == 0


def test_a_truncated_file_storage_peer_fails_the_consumer(env, tmp_path):
    """the same rule on the file:// storage-peer path: the consumer's transfer factory refuses the peer (a read fault of
    that producer), it is never handed header + partial payload as bytes"""
    cm, slots, lib, d = env
    data = os.urandom(100_000)
    p = tmp_path / "trunc.lzyb"
    hdr = bytearray(128)
    hdr[:8] = b"LZYB200\0"
    struct.pack_into("<IIQII", hdr, 8, N.lib.lzyb_abi_version(), N.DT_U8, len(data), 1, 1)
    struct.pack_into("<Q", hdr, 32, len(data))
    p.write_bytes(bytes(hdr) + data[:60_000])
    chan = cm.get_or_create("u", "wf", "exec-trunc", producer_uri=f"file://{p}")
    cons = slots.context("exec-trunc", "c")
    cons.add_slot("/t/in", chan, DIR_INPUT)
    with pytest.raises(N.LzybError, match="No more producers|Cannot get data|corrupt frame file"):
        cons.before_execution(10000)
    cons.close()
