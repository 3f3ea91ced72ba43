"""The C++ slot state machines (InputSlot / OutputSlot / SlotsExecutionContext / SlotsService) against
the reference's own slot tests, run over the test-only in-memory backend installed at the backend
seam — the role SlotsTest.InMemBackend plays in
lzy/slots/src/test/java/ai/lzy/slots/SlotsTest.java (testSimple :130-187, testConnect :216-256,
testInputRestart :339-393, testReadFromStorage / testWriteToStorage :395-462, testExecutionContext
:524-626, testLargeData :628-693, testRestoreReadFromOffset :803-864) and
lzy/channel-manager/src/test/java/ai/lzy/channelmanager/SlotsIntegrationTest.java (testStorageToConsumer
:119-163, testProducerToStorage :165-210, testMultiConsumer :212-297).  The channel manager is the
real in-memory one (the reference scripts a mock; a real matcher is a stronger test)."""
import os
import threading
import time

import numpy as np
import pytest

from lzy_b200 import _host as H
from lzy_b200 import _native as N
from lzy_b200._host import (DIR_INPUT, DIR_OUTPUT, ROLE_CONSUMER, ROLE_PRODUCER, ChannelManager, GrpcStatusError, Peer,
                            Slots)
from tests.support import hostmem


@pytest.fixture
def env():
    lib = hostmem.install()
    cm = ChannelManager(seed=3)
    H.pump_start(cm)
    slots = Slots(cm, 0, "lzyb://worker-0")
    yield cm, slots, lib
    slots.close()
    H.pump_stop(cm)
    cm.close()
    hostmem.uninstall()
    assert lib.hostmem_live_slots() == 0, "a slot frame leaked"


def produce(ctx, name, data: bytes):
    h = ctx.output_open(name, len(data))
    hostmem.write(h, data)
    ctx.output_done(name)
    return h


def test_simple_producer_then_consumer(env):
    """testSimple: 'Hello' travels output slot → input slot; the input re-publishes as '<id>-out'."""
    cm, slots, _ = env
    chan = cm.get_or_create("u", "wf", "exec", consumer_uri="mem://simple")
    out = slots.context("exec", "task-a")
    out.add_slot("2", chan, DIR_OUTPUT)
    out.before_execution()
    produce(out, "2", b"Hello")
    out.after_execution(5000)
    assert H.blob_get("mem://simple") == b"Hello"  # bound with a storage consumer → uploaded

    inp = slots.context("exec", "task-b")
    inp.add_slot("1", chan, DIR_INPUT)
    inp.before_execution(5000)
    assert hostmem.peek(inp.input_handle("1"), 5) == b"Hello"
    inp.after_execution(5000)  # waits for the "1-out" companion to be bound
    # the consumer peer is unbound by the input slot's own thread once it is done (InputSlot.clear(), InputSlot.java
    # :195-221) — not ordered against afterExecution(), in the reference either
    deadline = time.time() + 5
    while cm.channels_status("exec", [chan])[0]["consumers"] and time.time() < deadline:
        time.sleep(0.002)
    st = cm.channels_status("exec", [chan])[0]
    producers = sorted(p["peer_id"] for p in st["producers"] if p["kind"] == "slot")
    assert producers == ["1-out", "2"] and st["consumers"] == []
    # legacy Read on the companion and on the original producer, NOT_FOUND / UNIMPLEMENTED mapping
    assert slots.read("1-out", 0) == b"Hello" and slots.read("1-out", 5) is None
    assert slots.read("2", 1, 3) == b"ell"
    with pytest.raises(GrpcStatusError) as e:
        slots.read("nope", 0)
    assert e.value.status == "NOT_FOUND"
    inp.close()
    out.close()
    H.blob_delete("mem://simple")


def test_connect_early_consumer_gets_start_transfer(env):
    """testConnect / ApiTest.testEarlyConsumer: consumer binds first, waits for its peer, the producer's
    bind pushes StartTransfer."""
    cm, slots, _ = env
    chan = cm.get_or_create("u", "wf", "exec", consumer_uri="mem://early")
    inp = slots.context("exec", "task-b")
    inp.add_slot("in", chan, DIR_INPUT)
    inp.prefetch()
    deadline = time.time() + 5
    while inp.input_state("in") != H.IS_WAITING_FOR_PEER and time.time() < deadline:
        time.sleep(0.005)
    assert inp.input_state("in") == H.IS_WAITING_FOR_PEER
    # Read on an input slot is UNIMPLEMENTED; StartTransfer on an output slot is INTERNAL
    with pytest.raises(GrpcStatusError) as e:
        slots.read("in", 0)
    assert e.value.status == "UNIMPLEMENTED"
    out = slots.context("exec", "task-a")
    out.add_slot("out", chan, DIR_OUTPUT)
    out.before_execution()
    produce(out, "out", b"World")
    out.after_execution(5000)
    inp.before_execution(5000)
    assert hostmem.peek(inp.input_handle("in"), 5) == b"World"
    with pytest.raises(GrpcStatusError) as e:
        slots.start_transfer("out", Peer("x", H.PEER_SLOT, "u"), "t")
    assert e.value.status == "INTERNAL"
    with pytest.raises(GrpcStatusError) as e:
        slots.start_transfer("in", Peer("x", H.PEER_SLOT, "u"), "t")  # unregistered once it had a peer
    assert e.value.status == "NOT_FOUND"
    inp.close()
    out.close()
    H.blob_delete("mem://early")


def test_read_from_storage_and_write_to_storage(env):
    """testReadFromStorage / SlotsIntegrationTest.testStorageToConsumer + testProducerToStorage."""
    cm, slots, _ = env
    H.blob_put("mem://pre-existing", b"test")
    chan = cm.get_or_create("u", "wf", "exec", producer_uri="mem://pre-existing")
    inp = slots.context("exec", "t")
    inp.add_slot("in", chan, DIR_INPUT)
    inp.before_execution(5000)
    assert hostmem.peek(inp.input_handle("in"), 4) == b"test"
    inp.close()

    chan2 = cm.get_or_create("u", "wf", "exec", consumer_uri="mem://result")
    out = slots.context("exec", "t2")
    out.add_slot("out", chan2, DIR_OUTPUT)
    out.before_execution()
    produce(out, "out", b"result-bytes")
    out.after_execution(5000)
    assert H.blob_get("mem://result") == b"result-bytes"
    st = cm.channels_status("exec", [chan2])[0]
    assert len(st["producers"]) == 2 and st["consumers"] == []  # storage promoted to BACKUP producer
    out.close()
    H.blob_delete("mem://pre-existing")
    H.blob_delete("mem://result")


def test_execution_context_in_out_and_republish(env):
    """testExecutionContext: storage → in ('Hello'), out ('World') → storage, the delivered input is
    re-published and itself uploaded where its channel has a storage consumer."""
    cm, slots, _ = env
    H.blob_put("mem://key1", b"Hello")
    chan_in = cm.get_or_create("u", "wf", "exec", producer_uri="mem://key1")
    chan_out = cm.get_or_create("u", "wf", "exec", consumer_uri="mem://key2")
    ctx = slots.context("exec", "tid")
    ctx.add_slot("a/b/in", chan_in, DIR_INPUT)
    ctx.add_slot("a/b/out", chan_out, DIR_OUTPUT)
    ctx.before_execution(5000)
    assert hostmem.peek(ctx.input_handle("a/b/in"), 5) == b"Hello"
    produce(ctx, "a/b/out", b"World")
    ctx.after_execution(5000)
    assert H.blob_get("mem://key2") == b"World"
    st = cm.channels_status("exec", [chan_in])[0]
    assert "a/b/in-out" in [p["peer_id"] for p in st["producers"]]
    ctx.close()
    for k in ("mem://key1", "mem://key2"):
        H.blob_delete(k)


def test_large_data_200mib_equal(env):
    """testLargeData: 200 × 1 MiB random bytes, delivered == sent."""
    cm, slots, _ = env
    data = np.random.default_rng(1).integers(0, 256, 200 << 20, dtype=np.uint8).tobytes()
    chan = cm.get_or_create("u", "wf", "exec", consumer_uri="mem://large")
    slots.set_chunk_bytes(2 << 20)  # the reference's chunk size
    out = slots.context("exec", "a")
    out.add_slot("2", chan, DIR_OUTPUT)
    out.before_execution()
    produce(out, "2", data)
    out.after_execution(20000)
    inp = slots.context("exec", "b")
    inp.add_slot("1", chan, DIR_INPUT)
    inp.before_execution(20000)
    assert hostmem.peek(inp.input_handle("1"), len(data)) == data
    inp.close()
    out.close()
    H.blob_delete("mem://large")


def test_large_read_from_storage_and_large_write_to_storage(env):
    """testLargeReadFromStorage / testLargeWriteToStorage (SlotsTest.java:695-801): 200 MiB through the storage peer
    in each direction, delivered == stored"""
    cm, slots, _ = env
    data = np.random.default_rng(2).integers(0, 256, 200 << 20, dtype=np.uint8).tobytes()
    H.blob_put("mem://large-in", data)
    chan = cm.get_or_create("u", "wf", "exec-ls", producer_uri="mem://large-in")
    inp = slots.context("exec-ls", "r")
    inp.add_slot("/ls/in", chan, DIR_INPUT)
    inp.before_execution(30000)
    assert hostmem.peek(inp.input_handle("/ls/in"), len(data)) == data
    inp.close()
    H.blob_delete("mem://large-in")

    chan2 = cm.get_or_create("u", "wf", "exec-ls", consumer_uri="mem://large-out")
    out = slots.context("exec-ls", "w")
    out.add_slot("/ls/out", chan2, DIR_OUTPUT)
    out.before_execution()
    produce(out, "/ls/out", data)
    out.after_execution(30000)
    assert H.blob_get("mem://large-out") == data
    out.close()
    H.blob_delete("mem://large-out")


def test_restore_read_from_offset(env):
    """testRestoreReadFromOffset: 2 MiB, the producer fails after 1 MiB; the consumer reports
    TransferFailed, gets a peer again and resumes at the delivered offset."""
    cm, slots, lib = env
    data = os.urandom(2 << 20)
    chan = cm.get_or_create("u", "wf", "exec", consumer_uri="mem://restore")
    slots.set_chunk_bytes(256 << 10)
    out = slots.context("exec", "a")
    out.add_slot("2", chan, DIR_OUTPUT)
    out.before_execution()
    h_out = produce(out, "2", data)
    out.after_execution(5000)
    H.blob_delete("mem://restore")  # leave the slot producer as the only usable source of truth for bytes
    H.blob_put("mem://restore", data)
    lib.hostmem_set_fail_after(h_out, 1 << 20)

    inp = slots.context("exec", "b")
    inp.add_slot("1", chan, DIR_INPUT)
    inp.prefetch()
    # the slot producer (priority 7) is tried first and fails after 1 MiB; failures walk its priority
    # down until the storage backup (5) wins and the download completes from there
    inp.before_execution(10000)
    assert hostmem.peek(inp.input_handle("1"), len(data)) == data
    assert cm.peer_priority(chan, "2") < H.PRIO_PRIMARY
    inp.close()
    out.close()
    H.blob_delete("mem://restore")


def test_input_restart_on_corrupt_peer(env):
    """testInputRestart shape + digest verification: a peer that serves wrong bytes is a ReadException;
    the consumer fails over to another holder of the same data."""
    cm, slots, lib = env
    data = os.urandom(300_000)
    chan = cm.get_or_create("u", "wf", "exec", consumer_uri="mem://backup")
    out = slots.context("exec", "a")
    out.add_slot("p", chan, DIR_OUTPUT)
    out.before_execution()
    h = produce(out, "p", data)
    out.after_execution(5000)  # uploaded to the storage backup as well
    lib.hostmem_corrupt(h, 1234)  # the slot peer now serves one wrong byte
    inp = slots.context("exec", "b")
    inp.add_slot("c", chan, DIR_INPUT)
    inp.before_execution(10000)
    assert hostmem.peek(inp.input_handle("c"), len(data)) == data
    inp.close()
    out.close()
    H.blob_delete("mem://backup")


def test_multi_consumer_fanout(env):
    """SlotsIntegrationTest.testMultiConsumer: one producer, three consumers, all equal 'test'."""
    cm, slots, _ = env
    chan = cm.get_or_create("u", "wf", "exec", consumer_uri="mem://multi")
    ctxs = []
    for i in range(3):
        c = slots.context("exec", f"consumer-{i}")
        c.add_slot(f"in{i}", chan, DIR_INPUT)
        c.prefetch()
        ctxs.append(c)
    out = slots.context("exec", "producer")
    out.add_slot("out", chan, DIR_OUTPUT)
    out.before_execution()
    produce(out, "out", b"test")
    out.after_execution(5000)
    for i, c in enumerate(ctxs):
        c.before_execution(5000)
        assert hostmem.peek(c.input_handle(f"in{i}"), 4) == b"test"
    for c in ctxs:
        c.close()
    out.close()
    H.blob_delete("mem://multi")


def test_fail_open_is_a_local_fault_and_closes_the_context(env):
    """InMemBackend.failOpen: an IOException while opening the sink fails the slot (not the peer)."""
    cm, slots, lib = env
    H.blob_put("mem://x", b"payload")
    chan = cm.get_or_create("u", "wf", "exec", producer_uri="mem://x")
    lib.hostmem_set_fail_open(1)
    ctx = slots.context("exec", "t")
    ctx.add_slot("in", chan, DIR_INPUT)
    with pytest.raises(N.LzybError) as e:
        ctx.before_execution(5000)
    assert e.value.code == N.E_LOCAL and "Cannot open channel" in str(e.value)
    # like the reference, the failed consumer cannot unbind while its transfer is still ACTIVE
    # (Unbind → UNAVAILABLE, error ignored in InputSlot.clear, InputSlot.java:207-220)
    st = cm.channels_status("exec", [chan])[0]
    assert [c["peer_id"] for c in st["consumers"]] == ["in"]
    ctx.close()
    H.blob_delete("mem://x")


def test_fail_wait_completed_fails_after_execution(env):
    """InMemBackend.failWaitCompleted → afterExecution throws."""
    cm, slots, lib = env
    chan = cm.get_or_create("u", "wf", "exec", consumer_uri="mem://y")
    ctx = slots.context("exec", "t")
    ctx.add_slot("out", chan, DIR_OUTPUT)
    ctx.before_execution()
    lib.hostmem_set_fail_commit(1)
    produce(ctx, "out", b"zzz")
    with pytest.raises(N.LzybError) as e:
        ctx.after_execution(5000)
    assert "Cannot wait" in str(e.value)
    ctx.close()


def test_close_while_waiting_for_peer(env):
    """InputSlot.close from another thread while WAITING_FOR_PEER (InputSlot.java:110-117)."""
    cm, slots, _ = env
    chan = cm.get_or_create("u", "wf", "exec", consumer_uri="mem://never")
    ctx = slots.context("exec", "t")
    ctx.add_slot("in", chan, DIR_INPUT)
    ctx.prefetch()
    threading.Timer(0.1, ctx.close).start()
    with pytest.raises(N.LzybError):
        ctx.before_execution(5000)
    time.sleep(0.2)
    slot_consumers = [c for c in cm.channels_status("exec", [chan])[0]["consumers"] if c["kind"] == "slot"]
    assert slot_consumers == []  # clear() unbound it; only the channel's storage consumer is left


def test_only_producer_keeps_failing_destroys_the_channel(env):
    """A channel whose only producer (a storage peer whose object is gone) can never deliver: every
    TransferFailed walks the peer's priority down (SlotsService.java:249-253); once it is below zero
    there is no producer left, the channel manager destroys channel and workflow (:255-266) and the
    consumer's slot fails with that description."""
    cm, slots, lib = env
    chan = cm.get_or_create("u", "wf", "exec-np", producer_uri="mem://vanished")
    inp = slots.context("exec-np", "b")
    inp.add_slot("c", chan, DIR_INPUT)
    with pytest.raises(N.LzybError, match="No more producers in channel"):
        inp.before_execution(10000)
    assert cm.aborted("exec-np").startswith("No more producers")
    inp.close()


def test_stress_many_contexts_with_random_peer_failures(env):
    """40 producer → consumer pairs over 4 workers, a third of the producers fail part-way (the consumers
    fail over to the storage backup); everything is delivered bit-exact and no frame leaks."""
    import random
    cm, slots0, lib = env
    workers = [slots0] + [Slots(cm, 0, f"lzyb://worker-{i}") for i in range(1, 4)]
    rnd = random.Random(7)
    jobs = []
    try:
        def flaky(i, data):
            return i % 3 == 0 and len(data) > 64

        for i in range(40):
            data = os.urandom(rnd.choice([5, 4096, 70_001, 1 << 20]))
            chan = cm.get_or_create("u", "wf", "stress", consumer_uri=f"mem://stress/{i}")
            wp, wc = rnd.choice(workers), rnd.choice(workers)
            cons = wc.context("stress", f"c{i}")
            cons.add_slot(f"in{i}", chan, DIR_INPUT)
            if not flaky(i, data):
                cons.prefetch()  # early consumer
            prod = wp.context("stress", f"p{i}")
            prod.add_slot(f"out{i}", chan, DIR_OUTPUT)
            jobs.append((i, data, prod, cons, threading.Event()))

        errors = []

        def produce_all():
            try:
                for i, data, prod, cons, uploaded in jobs:
                    prod.before_execution()
                    h = prod.output_open(f"out{i}", len(data))
                    hostmem.write(h, data)
                    prod.output_done(f"out{i}")
                    if flaky(i, data):
                        # the storage backup must exist before the worker peer turns bad, otherwise the
                        # consumer exhausts the only producer's retries and the channel is destroyed
                        # (SlotsService.java:255-266) — test_only_producer_keeps_failing_destroys_the_channel
                        prod.after_execution(30000)
                        lib.hostmem_set_fail_after(h, (len(data) // 2) & ~15)
                    uploaded.set()
            except BaseException as e:  # surfaced in the main thread below
                errors.append(e)
                for j in jobs:
                    j[4].set()

        t = threading.Thread(target=produce_all)
        t.start()
        for i, data, prod, cons, uploaded in jobs:
            if flaky(i, data):
                assert uploaded.wait(30) and not errors, errors
            cons.before_execution(30000)
            assert hostmem.peek(cons.input_handle(f"in{i}"), len(data)) == data, f"job {i}"
        t.join()
        assert not errors, errors
        for i, data, prod, cons, _ in jobs:
            if not flaky(i, data):
                prod.after_execution(30000)
    finally:
        for _, _, prod, cons, _ in jobs:
            cons.close()
            prod.close()
        for i in range(40):
            H.blob_delete(f"mem://stress/{i}")
        for w in workers[1:]:
            w.close()


def test_transfer_log_records_bytes_time_source_and_failovers(env):
    """what the reference never measures: every input transfer leaves a record"""
    cm, slots, lib = env
    slots.transfer_log(clear=True)
    data = os.urandom(1 << 20)
    chan = cm.get_or_create("u", "wf", "exec-log", consumer_uri="mem://log/backup")
    out = slots.context("exec-log", "a")
    out.add_slot("/log/out", chan, DIR_OUTPUT)
    out.before_execution()
    h = produce(out, "/log/out", data)
    out.after_execution(5000)

    ok = slots.context("exec-log", "b")
    ok.add_slot("/log/in1", chan, DIR_INPUT)
    ok.before_execution(10000)
    lib.hostmem_set_fail_after(h, 4096)  # from now on the slot producer dies early: consumers end up on the backup
    failing = slots.context("exec-log", "c")
    failing.add_slot("/log/in2", chan, DIR_INPUT)
    failing.before_execution(10000)
    assert hostmem.peek(failing.input_handle("/log/in2"), len(data)) == data

    log = {r["slot_id"]: r for r in slots.transfer_log()}
    first, second = log["/log/in1"], log["/log/in2"]
    assert first["ok"] and first["bytes"] == len(data) and first["failovers"] == 0 and first["peer_kind"] == "slot"
    assert first["peer_id"] == "/log/out" and first["ns"] > 0 and first["error"] == "" and first["channel_id"] == chan
    assert second["ok"] and second["bytes"] == len(data) and second["failovers"] >= 1 and second["peer_kind"] == "storage"
    assert slots.transfer_log(clear=True) and slots.transfer_log() == []
    for c in (ok, failing, out):
        c.close()
    H.blob_delete("mem://log/backup")

    # a transfer that can never succeed is recorded as such
    gone = cm.get_or_create("u", "wf", "exec-log2", producer_uri="mem://log/vanished")
    lost = slots.context("exec-log2", "d")
    lost.add_slot("/log/in3", gone, DIR_INPUT)
    with pytest.raises(N.LzybError):
        lost.before_execution(10000)
    (rec,) = slots.transfer_log()
    assert not rec["ok"] and rec["slot_id"] == "/log/in3" and "No more producers" in rec["error"]
    lost.close()


def _os_threads() -> int:
    with open("/proc/self/status") as f:
        return next(int(line.split()[1]) for line in f if line.startswith("Threads:"))


@pytest.mark.skipif(os.environ.get("LZYB_SLOT_THREADS") == "fresh", reason="recycling is switched off for an A/B run")
def test_slot_threads_are_recycled(env):
    """every slot has a thread to block on, but a finished slot's thread serves the next slot: a long run of ops
    does not create a thread per slot (host/slots_runtime.cpp: SlotThreads)"""
    cm, slots, _ = env
    seen = []
    spawned_before = H.lib.lzyb_slots_threads_spawned()
    for k in range(150):
        chan = cm.get_or_create("u", "wf", "exec", consumer_uri=f"hbm://recycle/{k}")
        out = slots.context("exec", f"p{k}")
        out.add_slot(f"/p{k}/out", chan, DIR_OUTPUT)
        out.before_execution()
        produce(out, f"/p{k}/out", b"payload")
        out.after_execution(5000)
        inp = slots.context("exec", f"c{k}")
        inp.add_slot(f"/c{k}/in", chan, DIR_INPUT)
        inp.before_execution(5000)
        assert hostmem.peek(inp.input_handle(f"/c{k}/in"), 7) == b"payload"
        inp.close()
        out.close()
        seen.append(_os_threads())
    H.frame_store_clear("hbm://recycle/")
    assert max(seen[50:]) <= max(seen[:50]) + 2, f"thread count keeps growing: {seen[:5]} … {seen[-5:]}"
    assert H.lib.lzyb_slots_threads_spawned() - spawned_before <= 8, "300 slots ran on a handful of threads"


def test_cross_gpu_transfers_go_in_slices_same_gpu_ones_do_not():
    """lzyb_slots_set_peer_chunk_bytes: a transfer whose source frame is on another GPU is issued in slices (so that a
    kernel launched next to it gets the GPU at a slice boundary); a same-GPU transfer stays one call — it may be a
    hand-over of the whole frame, which only a whole-payload call can be"""
    lib = hostmem.install()
    cm = ChannelManager(seed=5)
    H.pump_start(cm)
    w0, w1 = Slots(cm, 0, "lzyb://worker-0"), Slots(cm, 1, "lzyb://worker-1")
    try:
        for w in (w0, w1):
            w.set_peer_chunk_bytes(256 << 10)
        data = os.urandom(1 << 20)
        for tag, consumer, want_calls in (("far", w1, 4), ("near", w0, 1)):
            chan = cm.get_or_create("u", "wf", f"exec-{tag}", consumer_uri=f"mem://slices-{tag}")
            out = w0.context(f"exec-{tag}", "a")
            out.add_slot("o", chan, DIR_OUTPUT)
            out.before_execution()
            produce(out, "o", data)
            out.after_execution(5000)
            before = lib.hostmem_moving_calls()
            inp = consumer.context(f"exec-{tag}", "b")
            inp.add_slot("i", chan, DIR_INPUT)
            inp.before_execution(5000)
            assert hostmem.peek(inp.input_handle("i"), len(data)) == data
            assert lib.hostmem_moving_calls() - before == want_calls, tag
            rec = [r for r in consumer.transfer_log(clear=True) if r["slot_id"] == "i"][-1]
            assert rec["ok"] and rec["bytes"] == len(data) and rec["ns_issue"] > 0
            inp.close()
            out.close()
            H.blob_delete(f"mem://slices-{tag}")
    finally:
        w0.close()
        w1.close()
        H.pump_stop(cm)
        cm.close()
        hostmem.uninstall()
    assert lib.hostmem_live_slots() == 0
