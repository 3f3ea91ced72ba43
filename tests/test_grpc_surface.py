"""The kept gRPC surface end to end over real sockets: a channel manager server and a slots server
(LzySlotsApi) in this process, reference-shaped clients, the in-memory test backend at the slot seam.
Scenarios: ApiTest.testSimple / testEarlyConsumer over the wire; a *legacy* consumer reading an output
slot with `Read` (2 MiB chunks + EOS, offset honoured, OutputSlot.java:238-265); status codes
(SlotsService.java:36,44,55,63; channel-manager SlotsService.java:184-187,219-222)."""
import os
import time

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


def test_deferred_start_transfer_is_delivered_over_the_wire():
    """A channel manager without the in-process pump: its START_TRANSFER actions go out as LzySlotsApi.StartTransfer
    to the consumer's worker (`RemoteActionPump`, ActionScheduler.startTransfer :61-106).  Early consumer on one
    worker, producer on another, each worker behind its own slots server."""
    from lzy_b200.grpc_api import RemoteActionPump
    hostmem.install()
    cm = H.ChannelManager(seed=3)
    servers, workers = [], []
    pump = None
    try:
        for _ in range(2):
            # a worker's URL is what it puts into BindRequest.peer_url: here directly its gRPC target
            probe, port = serve("127.0.0.1:0")  # find a free port, then bind the slots server to it
            probe.stop(0)
            w = H.Slots(cm, 0, f"127.0.0.1:{port}")
            server, bound = serve(f"127.0.0.1:{port}", slots=w)
            assert bound == port
            servers.append(server)
            workers.append(w)
        pump = RemoteActionPump(cm).start()
        data = os.urandom(300_000)
        chan = cm.get_or_create("u", "wf", "exec-wire", consumer_uri="mem://wire/out")
        cons = workers[0].context("exec-wire", "c")
        cons.add_slot("/wire/in", chan, H.DIR_INPUT)
        cons.prefetch()  # binds first: no producer yet, so it waits for StartTransfer
        prod = workers[1].context("exec-wire", "p")
        prod.add_slot("/wire/out", chan, H.DIR_OUTPUT)
        prod.before_execution()
        h = prod.output_open("/wire/out", len(data))
        hostmem.write(h, data)
        prod.output_done("/wire/out")
        cons.before_execution(10000)
        assert hostmem.peek(cons.input_handle("/wire/in"), len(data)) == data
        prod.after_execution(10000)
        deadline = time.time() + 5  # the download can finish before the StartTransfer RPC itself has returned
        while pump.delivered == 0 and time.time() < deadline:
            time.sleep(0.005)
        assert pump.delivered == 1 and pump.failed == 0
        cons.close()
        prod.close()

        # a consumer whose worker is gone: delivery fails, the channel is dropped and the workflow must be aborted
        chan2 = cm.get_or_create("u", "wf", "exec-wire2", consumer_uri="mem://wire/out2")
        lost = H.Slots(cm, 0, "127.0.0.1:1")  # nobody listens there
        ghost = lost.context("exec-wire2", "c")
        ghost.add_slot("/wire/in2", chan2, H.DIR_INPUT)
        ghost.prefetch()
        prod2 = workers[1].context("exec-wire2", "p")
        prod2.add_slot("/wire/out2", chan2, H.DIR_OUTPUT)
        prod2.before_execution()
        h2 = prod2.output_open("/wire/out2", 4)
        hostmem.write(h2, b"data")
        prod2.output_done("/wire/out2")
        deadline = time.time() + 15
        while pump.failed == 0 and time.time() < deadline:
            time.sleep(0.01)
        assert pump.failed == 1 and cm.aborted("exec-wire2")
        ghost.close()
        prod2.close()
        lost.close()
    finally:
        if pump is not None:
            pump.stop()
        for s in servers:
            s.stop(0)
        for w in workers:
            w.close()
        for uri in ("mem://wire/out", "mem://wire/out2"):
            H.blob_delete(uri)
        cm.close()
        hostmem.uninstall()


def test_channel_manager_service_entry_point_serves_the_private_api():
    """`python -m lzy_b200.grpc_api channel-manager`: a separate process, no GPU"""
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    proc = subprocess.Popen([sys.executable, "-m", "lzy_b200.grpc_api", "channel-manager", "--port", "0", "--seed", "1"],
                            cwd=str(root), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    try:
        line = proc.stdout.readline()
        assert line.startswith("channel-manager listening on"), line
        target = line.strip().rsplit(" ", 1)[-1]
        priv, pub = ChannelManagerPrivateClient(target), ChannelManagerClient(target)
        chan = priv.call("GetOrCreate", M.GetOrCreateRequest(execution_id="e", workflow_name="wf", user_id="u",
                                                             producer=storage_peer("s3://bucket/x"))).channel_id
        resp = pub.call("Bind", M.BindRequest(execution_id="e", channel_id=chan, peer_id="c1", peer_url="w:1", role=1))
        assert resp.peer.storage_peer.storage_uri == "s3://bucket/x" and resp.transfer_id
        st = priv.call("Status", M.StatusRequest(channel_id=chan)).channel
        assert [p.peer_id for p in st.consumers] == ["c1"] and len(st.producers) == 1
        priv.close()
        pub.close()
    finally:
        proc.terminate()
        assert proc.wait(timeout=15) == 0


def test_workers_against_a_channel_manager_in_another_process():
    """The control plane split the reference runs with: the channel manager is its own process
    (`python -m lzy_b200.grpc_api channel-manager`), this process holds two workers whose slot state machines reach
    it through LzyChannelManager RPCs (`remote_cm.connect`) and which it reaches through LzySlotsApi.StartTransfer.
    Early consumer, then a producer that dies half-way with failover to the storage backup."""
    import subprocess
    import sys
    from pathlib import Path

    from lzy_b200.grpc_api import remote_cm
    root = Path(__file__).resolve().parent.parent
    proc = subprocess.Popen([sys.executable, "-m", "lzy_b200.grpc_api", "channel-manager", "--port", "0", "--seed", "2"],
                            cwd=str(root), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    lib = hostmem.install()
    servers, workers, cm = [], [], None
    try:
        target = proc.stdout.readline().strip().rsplit(" ", 1)[-1]
        priv = ChannelManagerPrivateClient(target)
        cm = remote_cm.connect(target)
        for _ in range(2):
            probe, port = serve("127.0.0.1:0")
            probe.stop(0)
            w = H.Slots(cm, 0, f"127.0.0.1:{port}")
            server, _bound = serve(f"127.0.0.1:{port}", slots=w)
            servers.append(server)
            workers.append(w)

        def channel(uri):
            return priv.call("GetOrCreate", M.GetOrCreateRequest(execution_id="exec-split", workflow_name="wf", user_id="u",
                                                                 consumer=storage_peer(uri))).channel_id

        # 1. early consumer on worker 0, producer on worker 1
        data = os.urandom(200_000)
        chan = channel("mem://split/a")
        cons = workers[0].context("exec-split", "c")
        cons.add_slot("/split/in", chan, H.DIR_INPUT)
        cons.prefetch()
        prod = workers[1].context("exec-split", "p")
        prod.add_slot("/split/out", chan, H.DIR_OUTPUT)
        prod.before_execution()
        h = prod.output_open("/split/out", len(data))
        hostmem.write(h, data)
        prod.output_done("/split/out")
        cons.before_execution(15000)
        assert hostmem.peek(cons.input_handle("/split/in"), len(data)) == data
        prod.after_execution(15000)  # storage upload reported with TransferCompleted over the wire
        st = priv.call("Status", M.StatusRequest(channel_id=chan)).channel
        assert len(st.producers) >= 2  # the slot producer and the promoted storage backup

        # 2. the slot producer turns bad: TransferFailed over the wire walks its priority down, the backup serves
        lib.hostmem_set_fail_after(h, 4096)
        late = workers[0].context("exec-split", "late")
        late.add_slot("/split/in2", chan, H.DIR_INPUT)
        late.before_execution(15000)
        assert hostmem.peek(late.input_handle("/split/in2"), len(data)) == data
        rec = [r for r in workers[0].transfer_log() if r["slot_id"] == "/split/in2"][0]
        assert rec["failovers"] >= 1 and rec["peer_kind"] == "storage"

        # 3. status codes survive the wire: Unbind of a peer nobody knows is NOT_FOUND to the state machine as well
        for c in (late, cons, prod):
            c.close()
        priv.close()
    finally:
        for s in servers:
            s.stop(0)
        for w in workers:
            w.close()
        remote_cm.disconnect_all()
        H.blob_delete("mem://split/a")
        hostmem.uninstall()
        proc.terminate()
        proc.wait(timeout=15)


def test_consumer_pulls_from_a_producer_in_another_process(tmp_path):
    """Three processes: channel manager, a producer worker, and (here) a consumer worker.  The consumer's Bind names
    a slot peer that is not in its own registry, so the transfer factory asks the peer's worker for the frame's
    descriptor (LzybPeerApi.Describe → lzyb_slots_describe) and maps it (the backend's peer_import: CUDA IPC on a GPU
    box, POSIX shared memory in the in-memory test backend), then pulls and verifies as from a local peer."""
    import subprocess
    import sys
    from pathlib import Path

    import oracle
    from lzy_b200.grpc_api import peers, remote_cm
    root = Path(__file__).resolve().parent.parent
    nbytes, seed = (3 << 20) + 17, 11
    cm_proc = subprocess.Popen([sys.executable, "-m", "lzy_b200.grpc_api", "channel-manager", "--port", "0", "--seed", "4"],
                               cwd=str(root), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    producer = None
    hostmem.install()
    peers.install()
    slots = server = None
    try:
        cm_target = cm_proc.stdout.readline().strip().rsplit(" ", 1)[-1]
        priv = ChannelManagerPrivateClient(cm_target)
        backup = tmp_path / "store" / "mp.lzyb"  # a file:// storage consumer: the producer process uploads its output there
        chan = priv.call("GetOrCreate", M.GetOrCreateRequest(execution_id="exec-mp", workflow_name="wf", user_id="u",
                                                             consumer=storage_peer(f"file://{backup}"))).channel_id
        producer = subprocess.Popen([sys.executable, str(root / "tests/support/remote_worker.py"), cm_target, chan,
                                     str(nbytes), str(seed)], cwd=str(root), stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                    stderr=subprocess.PIPE, text=True)
        line = producer.stdout.readline()
        assert line.startswith("READY"), line + producer.stderr.read()
        producer_url = line.split()[1]

        cm = remote_cm.connect(cm_target)
        probe, port = serve("127.0.0.1:0")
        probe.stop(0)
        slots = H.Slots(cm, 0, f"127.0.0.1:{port}")
        server, _ = serve(f"127.0.0.1:{port}", slots=slots)
        before = peers.resolved
        cons = slots.context("exec-mp", "consumer")
        cons.add_slot("/local/in", chan, H.DIR_INPUT)
        cons.before_execution(20000)
        want = oracle.pattern(nbytes, seed).tobytes()
        assert hostmem.peek(cons.input_handle("/local/in"), nbytes) == want
        (rec,) = [r for r in slots.transfer_log() if r["slot_id"] == "/local/in"]
        assert rec["ok"] and rec["peer_kind"] == "slot" and rec["peer_id"] == "/remote/out" and rec["bytes"] == nbytes
        assert peers.resolved == before + 1  # one descriptor fetched from the other worker
        st = priv.call("Status", M.StatusRequest(channel_id=chan)).channel
        assert any(p.slot_peer.peer_url == producer_url for p in st.producers)
        cons.close()

        # the producer's process goes away; its output survives as the channel's file:// backup, which a late consumer
        # in this process reads after the dead slot peer has used up its retries
        assert backup.exists()
        producer.stdin.write("bye\n")
        producer.stdin.flush()
        producer.wait(timeout=20)
        producer = None
        late = slots.context("exec-mp", "late")
        late.add_slot("/local/in2", chan, H.DIR_INPUT)
        late.before_execution(30000)
        assert hostmem.peek(late.input_handle("/local/in2"), nbytes) == want
        (rec2,) = [r for r in slots.transfer_log() if r["slot_id"] == "/local/in2"]
        assert rec2["ok"] and rec2["peer_kind"] == "storage"
        late.close()
        priv.close()
    finally:
        if producer is not None:
            try:
                producer.stdin.write("bye\n")
                producer.stdin.flush()
                producer.wait(timeout=20)
            except Exception:  # noqa: BLE001
                producer.kill()
        if server is not None:
            server.stop(0)
        if slots is not None:
            slots.close()
        peers.uninstall()
        remote_cm.disconnect_all()
        hostmem.uninstall()
        cm_proc.terminate()
        cm_proc.wait(timeout=15)


def test_worker_entry_point_refuses_to_start_without_a_gpu():
    import subprocess
    import sys
    from pathlib import Path

    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for GPU-less boxes")
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, "-m", "lzy_b200.grpc_api", "worker", "--gpu", "0", "--port", "18999",
                        "--channel-manager", "127.0.0.1:1"], cwd=str(root), capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "needs a CUDA device" in r.stdout


def test_unreachable_remote_peer_is_a_read_failure_and_the_backup_serves(stack):
    """a slot producer registered by a worker that has since died: resolving it fails, the consumer reports
    TransferFailed until the channel manager hands out the storage backup; Describe of an unknown slot is NOT_FOUND"""
    from lzy_b200.grpc_api import PeerApiClient, peers
    from lzy_b200.grpc_api.protos import DescribeRequest
    cm, slots, pub, priv, sapi = stack
    peers.install()
    try:
        data = os.urandom(50_000)
        H.blob_put("mem://ghost/backup", data)
        chan = cm.get_or_create("u", "wf", "exec-ghost", producer_uri="mem://ghost/backup")
        cm.bind(chan, H.ROLE_PRODUCER, "/ghost/out", "127.0.0.1:1", "exec-ghost")  # PRIMARY, but nobody is there
        cons = slots.context("exec-ghost", "c")
        cons.add_slot("/ghost/in", chan, H.DIR_INPUT)
        cons.before_execution(20000)
        assert hostmem.peek(cons.input_handle("/ghost/in"), len(data)) == data
        (rec,) = [r for r in slots.transfer_log() if r["slot_id"] == "/ghost/in"]
        assert rec["failovers"] >= 1 and rec["peer_kind"] == "storage"
        assert cm.peer_priority(chan, "/ghost/out") < H.PRIO_PRIMARY
        cons.close()

        client = PeerApiClient(sapi.target)
        with pytest.raises(grpc.RpcError) as e:
            client.call("Describe", DescribeRequest(peer_id="/no/such/slot"))
        assert e.value.code() == grpc.StatusCode.NOT_FOUND
        client.close()
    finally:
        peers.uninstall()
        H.blob_delete("mem://ghost/backup")


def test_tasks_on_worker_processes_exchange_data_across_processes(tmp_path):
    """The CPU rehearsal of tests/test_multiprocess_gpu.py: a channel manager process and two worker *processes*
    (in-memory backend), driven only through WorkerApi.Execute / LongRunningService.Get.  Task `make` runs on worker
    A; task `check` runs on worker B, pulls A's output across the process boundary (LzybPeerApi.Describe + the
    backend's import) and verifies it inside the op body."""
    import socket
    import subprocess
    import sys
    from pathlib import Path

    from lzy_b200.api.v1.startup import ProcessingRequest
    from lzy_b200.api.v1.utils.pickle import pickle
    from lzy_b200.grpc_api import LongRunningClient, WorkerApiClient
    from lzy_b200.grpc_api.protos import GetOperationRequest
    root = Path(__file__).resolve().parent.parent
    n = 300_000

    def make() -> bytes:
        return bytes(i % 251 for i in range(n))

    def check(b: bytes) -> int:
        assert b == bytes(i % 251 for i in range(n))
        return len(b)

    def free_port():
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            return s.getsockname()[1]

    def execute(target, task_id, op, args_paths, outs, slots, assign):
        wapi, ops = WorkerApiClient(target), LongRunningClient(target)
        try:
            request = ProcessingRequest({}, [], op, args_paths, {}, outs, (tuple, f"/{task_id}/exc"), lazy_arguments=False)
            req = M.ExecuteRequest(task_id=task_id, execution_id="exec-wp", workflow_name="wf", user_id="u")
            req.task_desc.operation.command = "python -u startup.py " + pickle(request)
            for name, direction in slots:
                req.task_desc.operation.slots.add(name=name, direction=direction)
            for name, chan in assign.items():
                req.task_desc.slot_assignments.add(slot_name=name, channel_id=chan)
            op_msg = wapi.call("Execute", req)
            deadline = time.time() + 60
            while not op_msg.done:
                assert time.time() < deadline
                time.sleep(0.01)
                op_msg = ops.call("Get", GetOperationRequest(operation_id=op_msg.id))
            resp = M.ExecuteResponse()
            assert op_msg.WhichOneof("result") == "response" and op_msg.response.Unpack(resp)
            return resp
        finally:
            wapi.close()
            ops.close()

    procs = []
    try:
        cm = subprocess.Popen([sys.executable, "-m", "lzy_b200.grpc_api", "channel-manager", "--port", "0", "--seed", "6"],
                              cwd=str(root), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        procs.append(cm)
        cm_target = cm.stdout.readline().strip().rsplit(" ", 1)[-1]
        targets = []
        for _ in range(2):
            port = free_port()
            w = subprocess.Popen([sys.executable, str(root / "tests/support/remote_worker_api.py"), cm_target, str(port)],
                                 cwd=str(root), stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            procs.append(w)
            line = w.stdout.readline()
            assert line.startswith("READY"), line + w.stderr.read()
            targets.append(f"127.0.0.1:{port}")
            init = WorkerApiClient(targets[-1])
            init.call("Init", M.InitRequest(user_id="u", workflow_name="wf"))
            init.close()
        priv = ChannelManagerPrivateClient(cm_target)

        def channel(name):
            return priv.call("GetOrCreate", M.GetOrCreateRequest(execution_id="exec-wp", workflow_name="wf", user_id="u",
                                                                 consumer=storage_peer(f"file://{tmp_path}/{name}"))).channel_id

        data = channel("make.ret")
        r = execute(targets[0], "make", make, [], [(bytes, "/make/ret")], [("/make/ret", 2), ("/make/exc", 2)],
                    {"/make/ret": data, "/make/exc": channel("make.exc")})
        assert (r.rc, r.description) == (0, "Success")
        r = execute(targets[1], "check", check, [(bytes, "/check/arg")], [(int, "/check/ret")],
                    [("/check/arg", 1), ("/check/ret", 2), ("/check/exc", 2)],
                    {"/check/arg": data, "/check/ret": channel("check.ret"), "/check/exc": channel("check.exc")})
        assert (r.rc, r.description) == (0, "Success")
        assert (tmp_path / "check.ret").exists() and (tmp_path / "make.ret").stat().st_size >= n
        priv.close()
    finally:
        for p in procs[1:]:
            try:
                p.stdin.write("bye\n")
                p.stdin.flush()
            except Exception:  # noqa: BLE001
                pass
        for p in procs[1:]:
            try:
                p.wait(timeout=20)
            except subprocess.TimeoutExpired:
                p.kill()
        procs[0].terminate()
        procs[0].wait(timeout=15)
