"""Conformance of the in-memory channel manager with the reference's RPC semantics.

Scenarios restate lzy/channel-manager/src/test/java/ai/lzy/channelmanager/ApiTest.java (testSimple
:104-133, testStorageConsumer :144-176, testEarlyConsumer :178-199, testPriorityAfterFail :201-231,
testGetExistingChannel :233-256, testFailedToLoadDataToStorage :258-283, testUnbind :285-314) and
the DAO behaviours of DbTest.java (:163-274), then a randomized differential run against the
oracle model (oracle/channel_rules.py)."""
import random

import pytest

from lzy_b200 import _host as H
from lzy_b200._host import ROLE_CONSUMER, ROLE_PRODUCER, ChannelManager, GrpcStatusError
from oracle.channel_rules import ChannelRules, Status

URL = "http://slots-api:1234"


@pytest.fixture
def cm():
    m = ChannelManager(seed=7)
    yield m
    m.close()


def test_simple(cm):
    chan = cm.get_or_create("user", "wf", "execId", producer_uri="s3://some-bucket")
    peer, tid = cm.bind(chan, ROLE_CONSUMER, "1", URL, "execId", "bind")
    assert peer.is_storage and peer.url == "s3://some-bucket"
    st = cm.channels_status("execId", [chan])
    assert len(st) == 1 and len(st[0]["consumers"]) == 1 and len(st[0]["producers"]) == 1
    assert st[0]["producers"][0]["peer_id"] == peer.peer_id and st[0]["consumers"][0]["peer_id"] == "1"
    cm.transfer_completed(chan, tid, "complete")


def test_storage_consumer(cm):
    chan = cm.get_or_create("user", "wf", "execId", consumer_uri="s3://some-bucket")
    peer, tid = cm.bind(chan, ROLE_PRODUCER, "1", URL, "execId", "bind-1")
    assert peer.is_storage and peer.url == "s3://some-bucket"
    cm.transfer_completed(chan, tid, "complete")
    st = cm.channels_status("execId", [chan])
    assert len(st[0]["consumers"]) == 0 and len(st[0]["producers"]) == 2  # storage reconnected as producer
    peer2, _ = cm.bind(chan, ROLE_CONSUMER, "2", URL, "execId", "bind-2")
    assert peer2.peer_id == "1"  # the slot producer (PRIMARY=7) beats the storage backup (5)


def test_early_consumer(cm):
    chan = cm.get_or_create("user", "wf", "execId", consumer_uri="s3://some-bucket")
    peer, tid = cm.bind(chan, ROLE_CONSUMER, "1", URL, "execId", "bind-1")
    assert peer is None and tid is None
    assert cm.poll_action() is None
    peer2, _ = cm.bind(chan, ROLE_PRODUCER, "2", URL, "execId", "bind-2")
    assert peer2.url == "s3://some-bucket"
    act = cm.poll_action(100)
    assert act["kind"] == H.ACT_START_TRANSFER and act["consumer_id"] == "1" and act["producer"].peer_id == "2"
    assert act["dial_url"] == URL  # the reference dials the producer's URL with slot_id = consumer id
    assert cm.transfer_state(chan, act["transfer_id"]) == H.TR_PENDING
    cm.action_result(chan, act["transfer_id"], True)
    assert cm.transfer_state(chan, act["transfer_id"]) == H.TR_ACTIVE


def test_priority_after_fail(cm):
    chan = cm.get_or_create("user", "wf", "execId", producer_uri="s3://some-bucket")
    cm.bind(chan, ROLE_PRODUCER, "1", URL, "execId", "bind-1")
    cm.bind(chan, ROLE_PRODUCER, "2", URL, "execId", "bind-2")
    peer3, tid3 = cm.bind(chan, ROLE_CONSUMER, "3", URL, "execId", "bind-3")
    assert peer3.peer_id in ("1", "2")
    new_peer, new_tid = cm.transfer_failed(chan, tid3, "Fail", "failed")
    assert new_peer.peer_id != peer3.peer_id and new_tid != tid3
    assert cm.peer_priority(chan, peer3.peer_id) == H.PRIO_PRIMARY - 1


def test_get_existing_channel(cm):
    a = cm.get_or_create("user", "wf", "execId", producer_uri="s3://some-bucket")
    b = cm.get_or_create("user", "wf", "execId", producer_uri="s3://some-bucket")
    assert a == b
    assert cm.get_or_create("user", "wf", "execId", consumer_uri="s3://some-bucket") != a
    assert cm.get_or_create("user", "wf", "other", producer_uri="s3://some-bucket") != a


def test_failed_to_load_data_to_storage(cm):
    chan = cm.get_or_create("user", "wf", "execId", consumer_uri="s3://some-bucket")
    _, tid = cm.bind(chan, ROLE_PRODUCER, "1", URL, "execId", "bind-1")
    with pytest.raises(GrpcStatusError) as e:
        cm.transfer_failed(chan, tid, "Fail", "failed")
    assert e.value.status == "INTERNAL"
    assert cm.aborted("execId") == "Uploading data to storage failed"
    act = cm.poll_action()
    assert act["kind"] == H.ACT_ABORT_WORKFLOW and act["execution_id"] == "execId"
    assert cm.channels_status("execId") == []


def test_unbind(cm):
    chan = cm.get_or_create("user", "wf", "execId", producer_uri="s3://some-bucket")
    cm.bind(chan, ROLE_PRODUCER, "1", URL, "execId", "bind")
    cm.unbind(chan, "1")
    st = cm.channels_status("execId", [chan])
    assert len(st) == 1 and len(st[0]["producers"]) == 1 and len(st[0]["consumers"]) == 0


def test_unbind_unavailable_while_transfer_active(cm):
    """SlotsService.unbind :156-187"""
    chan = cm.get_or_create("user", "wf", "execId", producer_uri="s3://b")
    cm.bind(chan, ROLE_PRODUCER, "p", URL, "execId", "b1")
    _, tid = cm.bind(chan, ROLE_CONSUMER, "c", URL, "execId", "b2")
    for peer in ("p", "c"):
        with pytest.raises(GrpcStatusError) as e:
            cm.unbind(chan, peer)
        assert e.value.status == "UNAVAILABLE"
    cm.transfer_completed(chan, tid, "done")
    cm.unbind(chan, "c")
    cm.unbind(chan, "p")


def test_transfer_state_machine_and_idempotency(cm):
    """TransferDao states + 'idempotent only for the same key' (SlotsService.java:216-223,319-327; DbTest :229-274)"""
    chan = cm.get_or_create("user", "wf", "execId", producer_uri="s3://b")
    cm.bind(chan, ROLE_PRODUCER, "p", URL, "execId", "b1")
    _, tid = cm.bind(chan, ROLE_CONSUMER, "c", URL, "execId", "b2")
    assert cm.transfer_state(chan, tid) == H.TR_ACTIVE
    cm.transfer_completed(chan, tid, "k1")
    cm.transfer_completed(chan, tid, "k1")  # replay with the same key: fine
    with pytest.raises(GrpcStatusError) as e:
        cm.transfer_completed(chan, tid, "k2")
    assert e.value.status == "FAILED_PRECONDITION"
    with pytest.raises(GrpcStatusError) as e:
        cm.transfer_failed(chan, tid, "late", "k3")
    assert e.value.status == "FAILED_PRECONDITION"
    with pytest.raises(GrpcStatusError) as e:
        cm.transfer_completed(chan, "no-such-transfer", "k4")
    assert e.value.status == "NOT_FOUND"


def test_bind_replay_same_key_is_idempotent_other_key_conflicts(cm):
    chan = cm.get_or_create("user", "wf", "execId", producer_uri="s3://b")
    a = cm.bind(chan, ROLE_CONSUMER, "c", URL, "execId", "same")
    b = cm.bind(chan, ROLE_CONSUMER, "c", URL, "execId", "same")
    assert a[1] == b[1] and a[0].peer_id == b[0].peer_id
    with pytest.raises(GrpcStatusError) as e:
        cm.bind(chan, ROLE_CONSUMER, "c", URL, "execId", "different")
    assert e.value.status == "INTERNAL"


def test_bind_validation(cm):
    chan = cm.get_or_create("user", "wf", "execId", producer_uri="s3://b")
    with pytest.raises(GrpcStatusError) as e:
        cm.bind(chan, 0, "x", URL, "execId", "k")
    assert e.value.status == "INVALID_ARGUMENT"
    with pytest.raises(GrpcStatusError) as e:
        cm.bind("channel-unknown", ROLE_CONSUMER, "x", URL, "execId", "k")
    assert e.value.status == "INTERNAL"
    with pytest.raises(GrpcStatusError) as e:
        cm.get_or_create("user", "wf", "execId")
    assert e.value.status == "INVALID_ARGUMENT"


def test_priority_exhaustion_falls_back_to_storage_then_aborts(cm):
    """PRIMARY=7 loses one per failure; the BACKUP storage producer (5) takes over once the slot
    producer drops below it; when nobody is left the channel is dropped and the workflow aborted
    (SlotsService.java:246-266, DbTest.peerDaoTestPriority)."""
    chan = cm.get_or_create("user", "wf", "execId", producer_uri="s3://b")
    cm.bind(chan, ROLE_PRODUCER, "p", URL, "execId", "bp")
    peer, tid = cm.bind(chan, ROLE_CONSUMER, "c", URL, "execId", "bc")
    assert peer.peer_id == "p"
    seen = []
    for i in range(40):
        try:
            peer, tid = cm.transfer_failed(chan, tid, f"fail {i}", f"f{i}")
        except GrpcStatusError as e:
            assert e.status == "INTERNAL"
            break
        seen.append(peer.peer_id)
    else:
        pytest.fail("never ran out of producers")
    # p: 7→6 (still chosen), 6→5 (tie with storage), … eventually both below 0
    assert seen[0] == "p" and any(s.startswith("storage-peer-") for s in seen)
    assert cm.aborted("execId").startswith("No more producers")


def test_multi_consumer_fanout_one_transfer_each(cm):
    """SlotsIntegrationTest.testMultiConsumer shape: 3 early consumers, one producer bind claims all."""
    chan = cm.get_or_create("user", "wf", "execId", consumer_uri="s3://b")
    for i in range(3):
        assert cm.bind(chan, ROLE_CONSUMER, f"c{i}", URL, "execId", f"bc{i}") == (None, None)
    storage, tid = cm.bind(chan, ROLE_PRODUCER, "p", URL, "execId", "bp")
    assert storage.is_storage
    acts = [cm.poll_action() for _ in range(3)]
    assert sorted(a["consumer_id"] for a in acts) == ["c0", "c1", "c2"]
    assert all(a["producer"].peer_id == "p" for a in acts) and cm.poll_action() is None
    assert len({a["transfer_id"] for a in acts} | {tid}) == 4


def test_restore_actions_requeues_pending(cm):
    chan = cm.get_or_create("user", "wf", "execId", consumer_uri="s3://b")
    cm.bind(chan, ROLE_CONSUMER, "c", URL, "execId", "bc")
    cm.bind(chan, ROLE_PRODUCER, "p", URL, "execId", "bp")
    first = cm.poll_action()
    assert cm.restore_actions() == 1  # still PENDING: ActionScheduler.restoreActions re-issues it
    again = cm.poll_action()
    assert again["transfer_id"] == first["transfer_id"]
    cm.action_result(chan, first["transfer_id"], True)
    assert cm.restore_actions() == 0


def test_start_transfer_failure_aborts_workflow(cm):
    chan = cm.get_or_create("user", "wf", "execId", consumer_uri="s3://b")
    cm.bind(chan, ROLE_CONSUMER, "c", URL, "execId", "bc")
    cm.bind(chan, ROLE_PRODUCER, "p", URL, "execId", "bp")
    act = cm.poll_action()
    cm.action_result(chan, act["transfer_id"], False, "Cannot connect.")
    assert cm.aborted("execId") == "Cannot connect."


def test_destroy_all(cm):
    a = cm.get_or_create("user", "wf", "e1", producer_uri="s3://a")
    cm.get_or_create("user", "wf", "e1", consumer_uri="s3://b")
    c = cm.get_or_create("user", "wf", "e2", producer_uri="s3://a")
    cm.destroy_all("e1")
    assert cm.channels_status("e1") == [] and len(cm.channels_status("e2")) == 1
    cm.destroy(c)
    assert cm.channels_status("e2") == []
    cm.destroy(a)  # idempotent


@pytest.mark.parametrize("seed", range(25))
def test_differential_against_oracle_model(seed):
    rnd = random.Random(seed)
    cm, model = ChannelManager(seed=seed), ChannelRules()
    try:
        kw = {"producer_uri": "s3://x"} if rnd.random() < 0.5 else {"consumer_uri": "s3://x"}
        chan = cm.get_or_create("u", "wf", "exec", **kw)
        mchan = model.get_or_create("u", "exec", **kw)
        active = {}  # real transfer id -> model transfer id
        live_peers, n = [], 0
        for step in range(60):
            op = rnd.choice(["bindc", "bindc", "bindp", "complete", "fail", "unbind", "start"])
            try:
                if op == "bindc":
                    n += 1
                    peer, tid = cm.bind(chan, ROLE_CONSUMER, f"c{n}", URL, "exec", f"k{step}")
                    mt = model.bind_consumer(mchan, f"c{n}", URL, None if peer is None else _m(peer, model, mchan))
                    live_peers.append(f"c{n}")
                    if tid:
                        active[tid] = mt
                elif op == "bindp":
                    n += 1
                    peer, tid = cm.bind(chan, ROLE_PRODUCER, f"p{n}", URL, "exec", f"k{step}")
                    ms, _ = model.bind_producer(mchan, f"p{n}", URL)
                    live_peers.append(f"p{n}")
                    assert (peer is None) == (ms is None)
                    if tid:
                        mtid = next(t.id for t in model.channels[mchan].transfers.values() if t.frm == f"p{n}" and t.to == ms)
                        active[tid] = mtid
                elif op == "start":
                    act = cm.poll_action()
                    if act and act["kind"] == H.ACT_START_TRANSFER:
                        cm.action_result(chan, act["transfer_id"], True)
                        model.mark_started(mchan, act["consumer_id"], act["producer"].peer_id)
                        mtid = next((t.id for t in model.channels[mchan].transfers.values()
                                     if t.frm == act["producer"].peer_id and t.to == act["consumer_id"]), None)
                        if mtid:
                            active[act["transfer_id"]] = mtid
                elif op in ("complete", "fail") and active:
                    tid = rnd.choice(sorted(active))
                    mt = active.pop(tid)
                    if op == "complete":
                        err = _both(lambda: cm.transfer_completed(chan, tid, f"k{step}"), lambda: model.transfer_completed(mchan, mt))
                    else:
                        real = _try(lambda: cm.transfer_failed(chan, tid, "x", f"k{step}"))
                        chosen = None if real[0] == "err" else _m(real[1][0], model, mchan)
                        mres = _try(lambda: model.transfer_failed(mchan, mt, chosen))
                        assert real[0] == mres[0], (real, mres)
                        if real[0] == "err":
                            assert real[1] == mres[1]
                        else:
                            active[real[1][1]] = mres[1]
                elif op == "unbind" and live_peers:
                    pid = rnd.choice(live_peers)
                    err = _both(lambda: cm.unbind(chan, pid), lambda: model.unbind(mchan, pid))
                    if err is None:
                        live_peers.remove(pid)
            except StopIteration:
                pass
            if mchan not in model.channels:
                assert cm.channels_status("exec", [chan]) == []
                assert cm.aborted("exec") == model.aborted["exec"]
                break
            st = cm.channels_status("exec", [chan])[0]
            prods, cons, _trs = model.summary(mchan)
            assert sorted((p["peer_id"] if p["kind"] == "slot" else "S", p["priority"]) for p in st["producers"]) == \
                sorted((i if not s else "S", pr) for i, pr, s in prods)
            assert sorted((c["peer_id"] if c["kind"] == "slot" else "S", c["connected"]) for c in st["consumers"]) == \
                sorted((i if not s else "S", conn) for i, conn, s in cons)
    finally:
        cm.close()


def _m(peer, model, mchan):
    """map a real peer to the model's id (storage peers have generated ids on both sides)"""
    if not peer.is_storage:
        return peer.peer_id
    return next(p.id for p in model.channels[mchan].peers.values() if p.storage)


def _try(fn):
    try:
        return "ok", fn()
    except GrpcStatusError as e:
        return "err", e.status
    except Status as e:
        return "err", e.code


def _both(real, model):
    a, b = _try(real), _try(model)
    a = a[1] if a[0] == "err" else None
    b = b[1] if b[0] == "err" else None
    assert a == b, (a, b)
    return a
