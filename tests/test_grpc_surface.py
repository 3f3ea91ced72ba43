"""The kept gRPC surface end to end over real sockets: a channel manager server and a slots server
(LzySlotsApi) in this process, reference-shaped clients, the in-memory test backend at the slot seam.
Scenarios: ApiTest.testSimple / testEarlyConsumer over the wire; a *legacy* consumer reading an output
slot with `Read` (2 MiB chunks + EOS, offset honoured, OutputSlot.java:238-265); status codes
(SlotsService.java:36,44,55,63; channel-manager SlotsService.java:184-187,219-222)."""
import os

import grpc
import pytest

from lzy_b200 import _host as H
from lzy_b200.grpc_api import (ChannelManagerClient, ChannelManagerPrivateClient, SlotsApiClient, messages, serve)
from tests.support import hostmem

M = messages


@pytest.fixture
def stack():
    hostmem.install()
    cm = H.ChannelManager(seed=11)
    H.pump_start(cm)
    slots = H.Slots(cm, 0, "grpc-test-worker")
    server, port = serve("127.0.0.1:0", slots=slots, cm=cm)
    target = f"127.0.0.1:{port}"
    pub, priv, sapi = ChannelManagerClient(target), ChannelManagerPrivateClient(target), SlotsApiClient(target)
    yield cm, slots, pub, priv, sapi
    for c in (pub, priv, sapi):
        c.close()
    server.stop(0)
    slots.close()
    H.pump_stop(cm)
    cm.close()
    hostmem.uninstall()


def storage_peer(uri):
    return M.PeerDescription.StoragePeer(storage_uri=uri)


def test_messages_round_trip_on_the_wire_format():
    req = M.BindRequest(execution_id="e", channel_id="c", peer_id="p", peer_url="u", role=2)
    raw = req.SerializeToString()
    # field numbers are the contract: 1=execution_id 2=channel_id 3=peer_id 4=peer_url 5=role(varint)
    assert raw == b"\x0a\x01e\x12\x01c\x1a\x01p\x22\x01u\x28\x02"
    chunk = M.ReadDataChunk(chunk=b"abc")
    assert chunk.SerializeToString() == b"\x12\x03abc" and M.ReadDataChunk(control=0).WhichOneof("kind") == "control"
    pd = M.PeerDescription(peer_id="x")
    pd.slot_peer.peer_url = "cuda+ipc://h:1?gpu=3"
    assert M.PeerDescription.FromString(pd.SerializeToString()).slot_peer.peer_url.endswith("gpu=3")
    ds = M.DataScheme(dataFormat="lzyb_tensor_frame")
    ds.metadata["lzy.device"] = "cuda"
    assert M.DataScheme.FromString(ds.SerializeToString()).metadata["lzy.device"] == "cuda"


def test_api_simple_over_grpc(stack):
    cm, slots, pub, priv, sapi = stack
    chan = priv.call("GetOrCreate", M.GetOrCreateRequest(execution_id="execId", workflow_name="wf", user_id="u",
                                                         producer=storage_peer("s3://some-bucket"))).channel_id
    again = priv.call("GetOrCreate", M.GetOrCreateRequest(execution_id="execId", workflow_name="wf", user_id="u",
                                                          producer=storage_peer("s3://some-bucket"))).channel_id
    assert chan == again
    resp = pub.call("Bind", M.BindRequest(execution_id="execId", channel_id=chan, peer_id="1", peer_url="http://x", role=1), "bind")
    assert resp.peer.storage_peer.storage_uri == "s3://some-bucket" and resp.transfer_id
    st = pub.call("GetChannelsStatus", M.GetChannelsStatusRequest(execution_id="execId", channel_ids=[chan]))
    assert len(st.channels) == 1 and len(st.channels[0].consumers) == 1 and len(st.channels[0].producers) == 1
    assert st.channels[0].consumers[0].peer_id == "1"
    pub.call("TransferCompleted", M.TransferCompletedRequest(transfer_id=resp.transfer_id, channel_id=chan), "complete")
    with pytest.raises(grpc.RpcError) as e:
        pub.call("TransferCompleted", M.TransferCompletedRequest(transfer_id=resp.transfer_id, channel_id=chan), "other-key")
    assert e.value.code() == grpc.StatusCode.FAILED_PRECONDITION
    with pytest.raises(grpc.RpcError) as e:
        pub.call("Bind", M.BindRequest(execution_id="execId", channel_id=chan, peer_id="z", peer_url="u", role=0))
    assert e.value.code() == grpc.StatusCode.INVALID_ARGUMENT
    lst = priv.call("List", M.ListRequest(execution_id="execId"))
    assert [c.channel_id for c in lst.statuses] == [chan]
    assert priv.call("Status", M.StatusRequest(channel_id=chan)).channel.channel_id == chan
    priv.call("DestroyAll", M.DestroyAllRequest(execution_id="execId", reason="done"))
    assert len(priv.call("List", M.ListRequest(execution_id="execId")).statuses) == 0


def test_unbind_unavailable_and_transfer_failed_over_grpc(stack):
    cm, slots, pub, priv, sapi = stack
    chan = priv.call("GetOrCreate", M.GetOrCreateRequest(execution_id="e", workflow_name="wf", user_id="u",
                                                         producer=storage_peer("s3://b"))).channel_id
    pub.call("Bind", M.BindRequest(execution_id="e", channel_id=chan, peer_id="p", peer_url="u", role=2), "b1")
    r = pub.call("Bind", M.BindRequest(execution_id="e", channel_id=chan, peer_id="c", peer_url="u", role=1), "b2")
    assert r.peer.peer_id == "p"
    with pytest.raises(grpc.RpcError) as e:
        pub.call("Unbind", M.UnbindRequest(peer_id="c", channel_id=chan))
    assert e.value.code() == grpc.StatusCode.UNAVAILABLE
    f = pub.call("TransferFailed", M.TransferFailedRequest(transfer_id=r.transfer_id, channel_id=chan, description="x"), "f1")
    assert f.new_transfer_id and f.new_peer.peer_id  # p (priority 6) or the storage backup (5): p still wins
    assert f.new_peer.peer_id == "p"


def test_legacy_consumer_reads_hbm_slot_over_grpc(stack):
    """A stock CPU worker can still pull an output slot with LzySlotsApi.Read."""
    cm, slots, pub, priv, sapi = stack
    data = os.urandom((5 << 20) + 12345)
    chan = cm.get_or_create("u", "wf", "e", consumer_uri="mem://legacy")
    ctx = slots.context("e", "task")
    ctx.add_slot("out", chan, H.DIR_OUTPUT)
    ctx.before_execution()
    h = ctx.output_open("out", len(data))
    hostmem.write(h, data)
    ctx.output_done("out")
    ctx.after_execution(5000)
    sizes = []
    got = bytearray()
    for chunk in sapi.call("Read", M.ReadDataRequest(peer_id="out", offset=0)):
        if chunk.WhichOneof("kind") == "control":
            assert chunk.control == 0  # EOS
            break
        sizes.append(len(chunk.chunk))
        got += chunk.chunk
    assert bytes(got) == data and sizes == [2 << 20, 2 << 20, (1 << 20) + 12345]
    assert sapi.read_all("out", offset=(1 << 20) + 7) == data[(1 << 20) + 7:]
    assert sapi.read_all("out", offset=len(data)) == b""
    with pytest.raises(grpc.RpcError) as e:
        sapi.read_all("missing")
    assert e.value.code() == grpc.StatusCode.NOT_FOUND
    ctx.close()
    H.blob_delete("mem://legacy")


def test_start_transfer_over_grpc_reaches_waiting_input_slot(stack):
    """testConnect: the channel manager's push arrives as a real StartTransfer RPC."""
    import time
    cm, slots, pub, priv, sapi = stack
    H.pump_stop(cm)  # this test plays the action scheduler itself, over the wire
    chan = cm.get_or_create("u", "wf", "e", consumer_uri="mem://st")
    inp = slots.context("e", "consumer")
    inp.add_slot("in", chan, H.DIR_INPUT)
    inp.prefetch()
    deadline = time.time() + 5
    while inp.input_state("in") != H.IS_WAITING_FOR_PEER and time.time() < deadline:
        time.sleep(0.005)
    with pytest.raises(grpc.RpcError) as e:
        sapi.read_all("in")
    assert e.value.code() == grpc.StatusCode.UNIMPLEMENTED
    out = slots.context("e", "producer")
    out.add_slot("out", chan, H.DIR_OUTPUT)
    out.before_execution()
    h = out.output_open("out", 5)
    hostmem.write(h, b"Hello")
    out.output_done("out")
    out.after_execution(5000)
    act = cm.poll_action(1000)
    assert act["kind"] == H.ACT_START_TRANSFER and act["consumer_id"] == "in"
    peer = M.PeerDescription(peer_id=act["producer"].peer_id)
    peer.slot_peer.peer_url = act["producer"].url
    cm.action_result(act["channel_id"], act["transfer_id"], True)  # activate first: see the pump in slots_runtime.cpp
    sapi.call("StartTransfer", M.StartTransferRequest(transfer_id=act["transfer_id"], slot_id="in", peer=peer))
    inp.before_execution(5000)
    assert hostmem.peek(inp.input_handle("in"), 5) == b"Hello"
    with pytest.raises(grpc.RpcError) as e:
        sapi.call("StartTransfer", M.StartTransferRequest(transfer_id="t", slot_id="out", peer=peer))
    assert e.value.code() == grpc.StatusCode.INTERNAL  # "Cannot start transfer in output slot"
    with pytest.raises(grpc.RpcError) as e:
        sapi.call("StartTransfer", M.StartTransferRequest(transfer_id="t", slot_id="nope", peer=peer))
    assert e.value.code() == grpc.StatusCode.NOT_FOUND
    inp.close()
    out.close()
    H.blob_delete("mem://st")
    H.pump_start(cm)
