"""The slot state machines against a *scripted* channel manager, call by call — the form the reference's own slot
tests take (lzy/slots/src/test/java/ai/lzy/slots/SlotsTest.java with its ``ChannelManagerMock``): testSimple
:130-187, testCannotBind :189-214, testFailOnInputBackend :258-290, testFailOnOutputBackend :292-337,
testInputRestart :339-393, testReadFromStorageFail :464-492, testWriteToStorageFail :494-522.
tests/test_slots_runtime.py runs the same machines against the real matcher; here every Bind / Unbind /
TransferCompleted / TransferFailed the slots make is seen, checked and answered by the test."""
import os
import time

import pytest

from lzy_b200 import _host as H
from lzy_b200 import _native as N
from lzy_b200._host import DIR_INPUT, DIR_OUTPUT, ROLE_CONSUMER, ROLE_PRODUCER, GrpcStatusError, Peer, Slots
from tests.support import hostmem
from tests.support.scripted_cm import ScriptedChannelManager

ADDRESS = "lzyb://worker-0"


@pytest.fixture
def env():
    lib = hostmem.install()
    cm = ScriptedChannelManager()
    slots = Slots(cm, 0, ADDRESS)
    yield cm, slots, lib
    slots.close()
    cm.close()
    lib.hostmem_set_fail_open(0)
    hostmem.uninstall()
    assert lib.hostmem_live_slots() == 0, "a slot frame leaked"


def produce(ctx, name, data: bytes):
    h = ctx.output_open(name, len(data))
    hostmem.write(h, data)
    ctx.output_done(name)
    return h


def slot_peer(peer_id: str) -> Peer:
    return Peer(peer_id, H.PEER_SLOT, ADDRESS)


def wait_for(predicate, timeout=5.0):
    deadline = time.time() + timeout
    while not predicate() and time.time() < deadline:
        time.sleep(0.005)
    assert predicate()


def test_simple_call_by_call(env):
    """testSimple: producer binds only once its data is complete; the consumer binds, pulls, reports
    TransferCompleted for exactly the transfer it was given, then binds its '<id>-out' companion as a producer;
    close() unbinds every peer that was bound."""
    cm, slots, _ = env
    out_bind = cm.on_bind("1")
    out = slots.context("exec", "task-a")
    out.add_slot("1", "chan", DIR_OUTPUT)
    out.before_execution()
    h = out.output_open("1", 5)
    hostmem.write(h, b"Hello")
    time.sleep(0.05)
    assert cm.made("bind") == [], "a producer is not visible before its data is complete (OutputSlot.java:161-165)"
    out.output_done("1")
    req = out_bind.get()
    assert (req["peer_id"], req["channel_id"], req["role"], req["peer_url"], req["execution_id"]) == \
        ("1", "chan", ROLE_PRODUCER, ADDRESS, "exec")
    assert req["cookie"] == cm.h and req["idempotency_key"]
    out_bind.complete()  # no storage consumer on this channel
    out.after_execution(5000)

    in_bind = cm.on_bind("2").complete(slot_peer("1"), "transfer-id")
    done = cm.on_transfer_completed("transfer-id")
    companion = cm.on_bind("2-out")
    inp = slots.context("exec", "task-b")
    inp.add_slot("2", "chan", DIR_INPUT)
    inp.prefetch()
    in_req = in_bind.get()
    assert (in_req["peer_id"], in_req["role"], in_req["channel_id"]) == ("2", ROLE_CONSUMER, "chan")
    assert done.get()["channel_id"] == "chan"
    done.complete()
    inp.before_execution(5000)
    assert hostmem.peek(inp.input_handle("2"), 5) == b"Hello"
    creq = companion.get()
    assert (creq["role"], creq["peer_url"]) == (ROLE_PRODUCER, ADDRESS)
    companion.complete()
    inp.after_execution(5000)
    assert cm.made("failed") == []
    inp.close()
    out.close()
    assert sorted(r["peer_id"] for r in cm.made("unbind")) == ["1", "2", "2-out"]


def test_cannot_bind(env):
    """testCannotBind: a Bind the channel manager refuses fails afterExecution() of an output slot and
    beforeExecution() of an input slot; close() still works."""
    cm, slots, _ = env
    out_bind = cm.on_bind("1")
    out = slots.context("exec", "a")
    out.add_slot("1", "chan", DIR_OUTPUT)
    out.before_execution()
    produce(out, "1", b"Hello")
    out_bind.get()
    out_bind.complete_exceptionally("Cannot bind")
    with pytest.raises(N.LzybError, match="Cannot bind"):
        out.after_execution(5000)

    in_bind = cm.on_bind("2")
    inp = slots.context("exec", "b")
    inp.add_slot("2", "chan", DIR_INPUT)
    inp.prefetch()
    in_bind.get()
    in_bind.complete_exceptionally("Cannot bind")
    with pytest.raises(N.LzybError, match="Cannot bind"):
        inp.before_execution(5000)
    inp.close()
    out.close()
    assert cm.made("completed") == [] and cm.made("failed") == []


def test_fail_on_input_backend(env):
    """testFailOnInputBackend: the sink cannot be opened when StartTransfer arrives → beforeExecution() fails; that
    is a local fault, not the peer's: no TransferFailed is charged to the producer."""
    cm, slots, lib = env
    inp = slots.context("exec", "b")
    inp.add_slot("1", "chan", DIR_INPUT)
    inp.prefetch()
    wait_for(lambda: inp.input_state("1") == H.IS_WAITING_FOR_PEER)
    out = slots.context("exec", "a")
    out.add_slot("2", "chan", DIR_OUTPUT)
    out.before_execution()
    produce(out, "2", b"Hello")
    out.after_execution(5000)
    lib.hostmem_set_fail_open(1)
    slots.start_transfer("1", slot_peer("2"), "transfer-id")
    with pytest.raises(N.LzybError):
        inp.before_execution(5000)
    lib.hostmem_set_fail_open(0)
    assert cm.made("failed") == [] and cm.made("completed") == []
    inp.close()
    out.close()


def test_fail_on_output_backend_without_another_producer(env):
    """testFailOnOutputBackend: the producer cannot serve its bytes → TransferFailed('transfer-id', chan, reason) →
    an empty answer (no other producer) → beforeExecution() fails."""
    cm, slots, lib = env
    inp = slots.context("exec", "b")
    inp.add_slot("1", "chan", DIR_INPUT)
    inp.prefetch()
    wait_for(lambda: inp.input_state("1") == H.IS_WAITING_FOR_PEER)
    out = slots.context("exec", "a")
    out.add_slot("2", "chan", DIR_OUTPUT)
    out.before_execution()
    h = produce(out, "2", b"Hello")
    out.after_execution(5000)
    lib.hostmem_set_fail_read(h, 1)
    failed = cm.on_transfer_failed("transfer-id")
    slots.start_transfer("1", slot_peer("2"), "transfer-id")
    req = failed.get()
    assert req["channel_id"] == "chan" and req["description"]
    failed.complete()
    with pytest.raises(N.LzybError):
        inp.before_execution(5000)
    assert cm.made("completed") == []
    inp.close()
    out.close()


def test_input_restart_with_the_peer_the_channel_manager_names(env):
    """testInputRestart: the first attempt fails, TransferFailed answers (new_peer, new_transfer_id), the slot
    starts over with them and reports TransferCompleted for the *new* transfer."""
    cm, slots, lib = env
    data = os.urandom(70_000)
    inp = slots.context("exec", "b")
    inp.add_slot("1", "chan", DIR_INPUT)
    inp.prefetch()
    wait_for(lambda: inp.input_state("1") == H.IS_WAITING_FOR_PEER)
    out = slots.context("exec", "a")
    out.add_slot("2", "chan", DIR_OUTPUT)
    out.before_execution()
    h = produce(out, "2", data)
    out.after_execution(5000)
    lib.hostmem_set_fail_read(h, 1)
    failed = cm.on_transfer_failed("transfer-id")
    slots.start_transfer("1", slot_peer("2"), "transfer-id")
    failed.get()
    lib.hostmem_set_fail_read(h, 0)
    done = cm.on_transfer_completed("transfer-id-2")
    failed.complete(slot_peer("2"), "transfer-id-2")
    done.get()
    done.complete()
    inp.before_execution(5000)
    assert hostmem.peek(inp.input_handle("1"), len(data)) == data
    assert [r["transfer_id"] for r in cm.made("failed")] == ["transfer-id"]
    assert [r["transfer_id"] for r in cm.made("completed")] == ["transfer-id-2"]
    inp.close()
    out.close()


def test_read_from_storage_fail(env):
    """testReadFromStorageFail: Bind answers a storage producer whose object does not exist → TransferFailed for that
    transfer → empty answer → beforeExecution() fails."""
    cm, slots, _ = env
    in_bind = cm.on_bind("1")
    failed = cm.on_transfer_failed("transfer-id")
    inp = slots.context("exec", "b")
    inp.add_slot("1", "chan", DIR_INPUT)
    inp.prefetch()
    in_bind.get()
    in_bind.complete(Peer("storage", H.PEER_STORAGE, "mem://lolkek/key1"), "transfer-id")
    assert failed.get()["channel_id"] == "chan"
    failed.complete()
    with pytest.raises(N.LzybError):
        inp.before_execution(5000)
    assert cm.made("completed") == []
    inp.close()


def test_write_to_storage_fail(env, tmp_path):
    """testWriteToStorageFail: Bind answers a storage consumer that cannot be written → TransferFailed for that
    transfer → afterExecution() fails."""
    cm, slots, _ = env
    blocker = tmp_path / "not-a-directory"
    blocker.write_bytes(b"x")
    out_bind = cm.on_bind("1")
    failed = cm.on_transfer_failed("transfer-id")
    out = slots.context("exec", "a")
    out.add_slot("1", "chan", DIR_OUTPUT)
    out.before_execution()
    produce(out, "1", b"Hello")
    out_bind.get()
    out_bind.complete(Peer("storage", H.PEER_STORAGE, f"file://{blocker}/key1"), "transfer-id")
    assert failed.get()["description"]
    failed.complete()
    with pytest.raises(N.LzybError):
        out.after_execution(5000)
    assert cm.made("completed") == []
    out.close()


def test_write_to_storage_reports_completed_before_after_execution_returns(env, tmp_path):
    """testWriteToStorage :430-462: the upload and its TransferCompleted are part of afterExecution()."""
    cm, slots, _ = env
    target = tmp_path / "bucket" / "key1"
    out_bind = cm.on_bind("1").complete(Peer("storage", H.PEER_STORAGE, f"file://{target}"), "transfer-id")
    out = slots.context("exec", "a")
    out.add_slot("1", "chan", DIR_OUTPUT)
    out.before_execution()
    produce(out, "1", b"Hello")
    out.after_execution(5000)
    assert out_bind.request["role"] == ROLE_PRODUCER
    assert [r["transfer_id"] for r in cm.made("completed")] == ["transfer-id"]
    assert target.read_bytes().endswith(b"Hello")
    out.close()


def test_unknown_slots_and_double_start_transfer(env):
    """SlotsService.java:59-66 and InputSlot.java:84-102: StartTransfer for a slot this worker does not have is
    NOT_FOUND; a second StartTransfer for a slot that already has its peer finds it gone as well."""
    cm, slots, _ = env
    with pytest.raises(GrpcStatusError) as e:
        slots.start_transfer("nobody", slot_peer("2"), "t")
    assert e.value.status == "NOT_FOUND"
    out = slots.context("exec", "a")
    out.add_slot("2", "chan", DIR_OUTPUT)
    out.before_execution()
    produce(out, "2", b"Hello")
    out.after_execution(5000)
    inp = slots.context("exec", "b")
    inp.add_slot("1", "chan", DIR_INPUT)
    inp.prefetch()
    wait_for(lambda: inp.input_state("1") == H.IS_WAITING_FOR_PEER)
    slots.start_transfer("1", slot_peer("2"), "t-1")
    inp.before_execution(5000)
    with pytest.raises(GrpcStatusError) as e:
        slots.start_transfer("1", slot_peer("2"), "t-2")
    assert e.value.status == "NOT_FOUND"
    assert [r["transfer_id"] for r in cm.made("completed")] == ["t-1"]
    inp.close()
    out.close()


def test_close_is_idempotent_and_ignores_unbind_errors(env):
    """OutputSlot.close / InputSlot.close (OutputSlot.java:90-133, InputSlot.java:195-221): Unbind failures are
    swallowed, the backing frames are released all the same (the fixture checks for leaks), a second close is a
    no-op and makes no second Unbind."""
    cm, slots, _ = env
    cm.on_unbind("2").complete_exceptionally("channel is gone", N.E_NOT_FOUND)
    cm.on_unbind("1").complete_exceptionally("try later", N.E_UNAVAILABLE)
    out = slots.context("exec", "a")
    out.add_slot("2", "chan", DIR_OUTPUT)
    out.before_execution()
    produce(out, "2", b"Hello")
    out.after_execution(5000)
    inp = slots.context("exec", "b")
    inp.add_slot("1", "chan", DIR_INPUT)
    inp.prefetch()
    wait_for(lambda: inp.input_state("1") == H.IS_WAITING_FOR_PEER)
    inp.close()  # while still waiting for a peer
    out.close()
    inp.close()
    out.close()
    assert sorted(r["peer_id"] for r in cm.made("unbind")) == ["1", "2"]
    with pytest.raises(GrpcStatusError) as e:
        slots.read("2", 0)
    assert e.value.status == "NOT_FOUND"  # unregistered from the slots server
