"""The box-local transport (csrc/host/local_rpc.cpp): the matcher's four calls and LzybPeerApi.Describe over unix sockets,
served and dialled by threads of the library — for this repository's own worker processes and channel manager on one box.
What must hold: the slot state machines see exactly what they see from the in-process matcher or over gRPC (same
scenarios, same status codes, same texts), both transports serve ONE matcher side by side, and a dead socket is
UNAVAILABLE, not a hang.  Scenarios after SlotsTest.testSimple / testConnect (lzy/slots/src/test/java/ai/lzy/slots/
SlotsTest.java:130-256) and ApiTest.testUnbind… (lzy/channel-manager/src/test/java/ai/lzy/channelmanager/test/ApiTest.java:285-314)."""
import os
import socket
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import pytest

from lzy_b200 import _host as H
from lzy_b200._host import DIR_INPUT, DIR_OUTPUT, ChannelManager, GrpcStatusError, LocalChannelManager, Slots
from tests.support import hostmem

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture
def sockdir():
    d = tempfile.mkdtemp(prefix="lzyb-", dir="/tmp")  # unix socket paths are short: not pytest's tmp_path
    yield d
    import shutil
    shutil.rmtree(d, ignore_errors=True)


@pytest.fixture
def env(sockdir):
    lib = hostmem.install()
    cm = ChannelManager(seed=3)
    H.pump_start(cm)
    cm.serve_local(f"{sockdir}/cm.sock")
    lcm = LocalChannelManager(f"{sockdir}/cm.sock")
    slots = Slots(lcm, 0, "lzyb://worker-local")
    yield cm, lcm, slots
    slots.close()
    lcm.close()
    H.pump_stop(cm)
    cm.close()
    hostmem.uninstall()
    assert lib.hostmem_live_slots() == 0, "a slot frame leaked"


def produce(ctx, name, data: bytes):
    h = ctx.output_open(name, len(data))
    hostmem.write(h, data)
    ctx.output_done(name)


def test_producer_then_consumer_through_the_socket(env):
    cm, _lcm, slots = env
    chan = cm.get_or_create("u", "wf", "exec", consumer_uri="mem://local-simple")
    out = slots.context("exec", "task-a")
    out.add_slot("2", chan, DIR_OUTPUT)
    out.before_execution()
    produce(out, "2", b"Hello")
    out.after_execution(5000)
    assert H.blob_get("mem://local-simple") == b"Hello"  # Bind answered with the storage consumer, upload, TransferCompleted

    inp = slots.context("exec", "task-b")
    inp.add_slot("1", chan, DIR_INPUT)
    inp.before_execution(5000)
    assert hostmem.peek(inp.input_handle("1"), 5) == b"Hello"
    inp.after_execution(5000)
    deadline = time.time() + 5
    while cm.channels_status("exec", [chan])[0]["consumers"] and time.time() < deadline:
        time.sleep(0.002)
    st = cm.channels_status("exec", [chan])[0]
    assert sorted(p["peer_id"] for p in st["producers"] if p["kind"] == "slot") == ["1-out", "2"] and st["consumers"] == []
    # every call of the two slots was answered by the socket's threads: the transfers are COMPLETED in the real matcher
    (rec,) = [r for r in slots.transfer_log() if r["slot_id"] == "1"]
    assert rec["ok"] and rec["bytes"] == 5 and rec["peer_id"] == "2"
    inp.close()
    out.close()
    H.blob_delete("mem://local-simple")


def test_early_consumer_is_started_by_the_matchers_push(sockdir):
    """the consumer binds first (empty answer over the socket) and waits; the producer's Bind makes the matcher push
    StartTransfer — which keeps travelling as LzySlotsApi.StartTransfer to the consumer's gRPC port, as between processes"""
    from lzy_b200.grpc_api import serve
    from lzy_b200.grpc_api.pump import RemoteActionPump
    lib = hostmem.install()
    cm = ChannelManager(seed=3)
    cm.serve_local(f"{sockdir}/cm.sock")
    pump = RemoteActionPump(cm).start()
    lcm = LocalChannelManager(f"{sockdir}/cm.sock")
    probe, port = serve("127.0.0.1:0")
    probe.stop(0)
    slots = Slots(lcm, 0, f"127.0.0.1:{port}")
    server, _ = serve(f"127.0.0.1:{port}", slots=slots)
    try:
        chan = cm.get_or_create("u", "wf", "exec", consumer_uri="mem://local-early")
        inp = slots.context("exec", "task-b")
        inp.add_slot("in", chan, DIR_INPUT)
        inp.prefetch()
        deadline = time.time() + 5
        while inp.input_state("in") != H.IS_WAITING_FOR_PEER and time.time() < deadline:
            time.sleep(0.005)
        assert inp.input_state("in") == H.IS_WAITING_FOR_PEER  # Bind came back empty over the socket
        out = slots.context("exec", "task-a")
        out.add_slot("out", chan, DIR_OUTPUT)
        out.before_execution()
        produce(out, "out", b"early bird")
        out.after_execution(5000)
        inp.before_execution(5000)
        assert hostmem.peek(inp.input_handle("in"), 10) == b"early bird"
        deadline = time.time() + 5  # the pump counts a delivery when StartTransfer has been answered: the consumer may be done first
        while pump.delivered < 1 and time.time() < deadline:
            time.sleep(0.005)
        assert pump.delivered == 1 and pump.failed == 0
        inp.close()
        out.close()
        H.blob_delete("mem://local-early")
    finally:
        server.stop(0)
        slots.close()
        lcm.close()
        pump.stop()
        cm.close()
        hostmem.uninstall()
    assert lib.hostmem_live_slots() == 0


def test_refusals_arrive_with_the_matchers_own_code_and_text(env):
    """the answer frame carries rc and lzyb_last_error of the serving thread: what a slot logs is what the matcher said"""
    cm, _lcm, slots = env
    direct = None
    try:
        cm.bind("no-such-channel", H.ROLE_CONSUMER, "p", "url", "exec")
    except GrpcStatusError as e:
        direct = (e.status, e.message)
    assert direct and direct[0] == "INTERNAL" and "no-such-channel" in direct[1]  # SlotsService.java:84-88
    ctx = slots.context("exec", "lost")
    ctx.add_slot("x", "no-such-channel", DIR_INPUT)
    with pytest.raises(H.N.LzybError) as e:
        ctx.before_execution(5000)
    assert direct[1] in str(e.value)  # "Bind failed: <the matcher's text>"
    ctx.close()


def test_garbage_on_the_socket_does_not_take_the_server_down(env, sockdir):
    """a connection that sends nonsense is answered UNIMPLEMENTED (unknown call), E_ARG (fields missing) or dropped (a frame
    longer than the limit, a truncated one) — and the next well-formed client is served as if nothing had happened"""
    import struct
    cm, _lcm, slots = env
    path = f"{sockdir}/cm.sock"

    def raw(payload: bytes, expect_answer=True):
        with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as sk:
            sk.settimeout(5)
            sk.connect(path)
            sk.sendall(payload)
            if not expect_answer:
                sk.shutdown(socket.SHUT_WR)
                return sk.recv(4)
            (n,) = struct.unpack("<I", sk.recv(4, socket.MSG_WAITALL))
            body = sk.recv(n, socket.MSG_WAITALL)
            rc, elen = struct.unpack_from("<iI", body, 0)
            return rc, body[8:8 + elen].decode()

    rc, text = raw(struct.pack("<II", 4, 99))                      # a call this socket does not serve
    assert rc == H.N.E_UNIMPLEMENTED and "99" in text
    rc, text = raw(struct.pack("<II", 4, 1))                       # Bind without its fields
    assert rc < 0 and "malformed Bind" in text
    rc, text = raw(struct.pack("<III", 8, 5, 0))                   # Describe asked of the matcher's socket
    assert rc == H.N.E_UNIMPLEMENTED
    assert raw(struct.pack("<I", 1 << 30), expect_answer=False) == b""   # a frame beyond the limit: dropped
    assert raw(struct.pack("<II", 400, 1) + b"x" * 10, expect_answer=False) == b""  # truncated: dropped when the peer gives up
    chan = cm.get_or_create("u", "wf", "exec", consumer_uri="mem://after-garbage")
    out = slots.context("exec", "fine")
    out.add_slot("o", chan, DIR_OUTPUT)
    out.before_execution()
    produce(out, "o", b"still here")
    out.after_execution(5000)
    assert H.blob_get("mem://after-garbage") == b"still here"
    out.close()
    H.blob_delete("mem://after-garbage")


def test_nobody_listening_is_unavailable_at_connect_time(sockdir):
    with pytest.raises(GrpcStatusError) as e:
        LocalChannelManager(f"{sockdir}/nobody.sock")
    assert e.value.status == "UNAVAILABLE"
    with pytest.raises(H.N.LzybError):
        LocalChannelManager("/tmp/" + "x" * 200)  # longer than sun_path


def test_a_matcher_that_went_away_fails_the_slot_instead_of_hanging_it(sockdir):
    lib = hostmem.install()
    cm = ChannelManager(seed=1)
    cm.serve_local(f"{sockdir}/gone.sock")
    lcm = LocalChannelManager(f"{sockdir}/gone.sock")
    slots = Slots(lcm, 0, "lzyb://worker-orphan")
    chan = cm.get_or_create("u", "wf", "exec", consumer_uri="mem://gone")
    try:
        cm.stop_local()
        ctx = slots.context("exec", "orphan")
        ctx.add_slot("x", chan, DIR_INPUT)
        t0 = time.time()
        with pytest.raises(H.N.LzybError) as e:
            ctx.before_execution(5000)
        assert time.time() - t0 < 4 and "cannot reach" in str(e.value)
        ctx.close()
        # ... and serving again on the same path brings the same client back (a connection is dialled per call when needed)
        cm.serve_local(f"{sockdir}/gone.sock")
        out = slots.context("exec", "back")
        out.add_slot("o", chan, DIR_OUTPUT)
        out.before_execution()
        produce(out, "o", b"again")
        out.after_execution(5000)
        assert H.blob_get("mem://gone") == b"again"
        out.close()
        H.blob_delete("mem://gone")
    finally:
        slots.close()
        lcm.close()
        cm.close()
        hostmem.uninstall()
    assert lib.hostmem_live_slots() == 0


def free_port() -> int:
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


@pytest.mark.parametrize("second_worker_local", [True, False], ids=["both-local", "second-over-grpc"])
def test_two_worker_processes_and_a_channel_manager_process(sockdir, second_worker_local):
    """The deployment shape: channel manager + two worker processes, tasks sent as WorkerApi.Execute.  `make` on worker 0,
    `check` on worker 1 pulls make's result from the other process — Bind and TransferCompleted over the channel manager's
    socket, the frame descriptor over worker 0's socket.  Second variant: worker 1 uses gRPC for both (a stock-shaped peer
    next to a local one): the matcher serves both transports at once and worker 0's URL still names its gRPC port."""
    from lzy_b200.api.v1.startup import ProcessingRequest
    from lzy_b200.api.v1.utils.pickle import pickle
    from lzy_b200.grpc_api import ChannelManagerPrivateClient, LongRunningClient, WorkerApiClient, messages as M
    n = (1 << 20) + 3

    def make() -> bytes:
        return bytes(range(251)) * (n // 251 + 1)

    def check(b: bytes) -> int:
        assert b == bytes(range(251)) * (n // 251 + 1)
        return len(b)

    procs, clients = [], []
    try:
        cm = subprocess.Popen([sys.executable, "-m", "lzy_b200.grpc_api", "channel-manager", "--port", "0", "--seed", "6",
                               "--local-socket", f"{sockdir}/cm.sock"], cwd=str(ROOT), stdout=subprocess.PIPE,
                              stderr=subprocess.DEVNULL, text=True)
        procs.append(cm)
        cm_target = cm.stdout.readline().strip().rsplit(" ", 1)[-1]
        for i in range(2):
            port = free_port()
            local = [f"{sockdir}/cm.sock", f"{sockdir}/w{i}.sock"] if (i == 0 or second_worker_local) else []
            w = subprocess.Popen([sys.executable, str(ROOT / "tests/support/remote_worker_api.py"), cm_target, str(port)] + local,
                                 cwd=str(ROOT), stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            procs.append(w)
            line = w.stdout.readline()
            assert line.startswith("READY"), line + w.stderr.read()
            clients.append((WorkerApiClient(f"127.0.0.1:{port}"), LongRunningClient(f"127.0.0.1:{port}")))
            clients[-1][0].call("Init", M.InitRequest(user_id="u", workflow_name="wf"))
        assert os.path.exists(f"{sockdir}/w0.sock") and os.path.exists(f"{sockdir}/w1.sock") == second_worker_local
        priv = ChannelManagerPrivateClient(cm_target)

        def channel(name):
            sp = M.PeerDescription.StoragePeer(storage_uri=f"file://{sockdir}/{name}")
            return priv.call("GetOrCreate", M.GetOrCreateRequest(execution_id="exec-lt", workflow_name="wf", user_id="u", consumer=sp)).channel_id

        def execute(i, task_id, fn, args_paths, outs, slots, assign):
            wapi, ops = clients[i]
            request = ProcessingRequest({}, [], fn, args_paths, {}, outs, (tuple, f"/{task_id}/exc"), lazy_arguments=False)
            req = M.ExecuteRequest(task_id=task_id, execution_id="exec-lt", workflow_name="wf", user_id="u")
            req.task_desc.operation.command = "python -u startup.py " + pickle(request)
            for name, direction in slots:
                req.task_desc.operation.slots.add(name=name, direction=direction)
            for name, chan in assign.items():
                req.task_desc.slot_assignments.add(slot_name=name, channel_id=chan)
            op = ops.wait(wapi.call("Execute", req).id, timeout=60.0)
            resp = M.ExecuteResponse()
            assert op.WhichOneof("result") == "response" and op.response.Unpack(resp)
            return resp

        data = channel("make.ret")
        r = execute(0, "make", make, [], [(bytes, "/make/ret")], [("/make/ret", 2), ("/make/exc", 2)],
                    {"/make/ret": data, "/make/exc": channel("make.exc")})
        assert (r.rc, r.description) == (0, "Success")
        r = execute(1, "check", check, [(bytes, "/check/arg")], [(int, "/check/ret")],
                    [("/check/arg", 1), ("/check/ret", 2), ("/check/exc", 2)],
                    {"/check/arg": data, "/check/ret": channel("check.ret"), "/check/exc": channel("check.exc")})
        assert (r.rc, r.description) == (0, "Success")
        # the matcher saw worker 0's slot with the URL that names its socket, and both transfers of `data` completed
        st = priv.call("Status", M.StatusRequest(channel_id=data)).channel
        urls = [p.slot_peer.peer_url for p in st.producers if p.WhichOneof("peer") == "slot_peer"]
        assert any(u.endswith(f"&uds={sockdir}/w0.sock") for u in urls), urls
        assert os.path.exists(f"{sockdir}/check.ret") and os.path.getsize(f"{sockdir}/make.ret") >= n
        priv.close()
    finally:
        for wapi, ops in clients:
            wapi.close()
            ops.close()
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


@pytest.mark.parametrize("local", [True, False], ids=["local-sockets", "grpc-only"])
def test_the_worker_service_itself_starts_serves_a_task_and_leaves_on_sigterm(sockdir, local):
    """`python -m lzy_b200.grpc_api worker` (through tests/support/worker_main_twin.py: the same main, in-memory backend) next
    to `python -m lzy_b200.grpc_api channel-manager`: the line it announces itself with, one task through WorkerApi.Execute
    answered on completion, SIGTERM → exit code 0 and no socket left behind"""
    import signal

    from lzy_b200.api.v1.startup import ProcessingRequest
    from lzy_b200.api.v1.utils.pickle import pickle
    from lzy_b200.grpc_api import ChannelManagerPrivateClient, WorkerApiClient, messages as M

    def forty_two() -> int:
        return 42

    cm = subprocess.Popen([sys.executable, "-m", "lzy_b200.grpc_api", "channel-manager", "--port", "0"] +
                          (["--local-socket", f"{sockdir}/cm.sock"] if local else []), cwd=str(ROOT), stdout=subprocess.PIPE,
                          stderr=subprocess.DEVNULL, text=True)
    w = None
    try:
        cm_target = cm.stdout.readline().strip().rsplit(" ", 1)[-1]
        port = free_port()
        w = subprocess.Popen([sys.executable, str(ROOT / "tests/support/worker_main_twin.py"), "worker", "--gpu", "0", "--port", str(port),
                              "--channel-manager", cm_target] +
                             (["--channel-manager-local", f"{sockdir}/cm.sock", "--local-socket", f"{sockdir}/w0.sock"] if local else []),
                             cwd=str(ROOT), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        line = w.stdout.readline()
        assert line.startswith("worker for GPU 0 listening on 127.0.0.1:"), line + w.stderr.read()
        assert line.strip().endswith(f"?gpu=0&uds={sockdir}/w0.sock" if local else "?gpu=0")
        wapi, priv = WorkerApiClient(f"127.0.0.1:{port}"), ChannelManagerPrivateClient(cm_target)
        wapi.call("Init", M.InitRequest(user_id="u", workflow_name="wf"))
        assert os.path.exists(f"{sockdir}/w0.sock") == local  # the Describe socket exists once the slots server does

        def channel(name):
            sp = M.PeerDescription.StoragePeer(storage_uri=f"file://{sockdir}/{name}")
            return priv.call("GetOrCreate", M.GetOrCreateRequest(execution_id="exec-m", workflow_name="wf", user_id="u", consumer=sp)).channel_id

        request = ProcessingRequest({}, [], forty_two, [], {}, [(int, "/t/ret")], (tuple, "/t/exc"), lazy_arguments=False)
        req = M.ExecuteRequest(task_id="t", execution_id="exec-m", workflow_name="wf", user_id="u")
        req.task_desc.operation.command = "python -u startup.py " + pickle(request)
        for name in ("/t/ret", "/t/exc"):
            req.task_desc.operation.slots.add(name=name, direction=2)
        req.task_desc.slot_assignments.add(slot_name="/t/ret", channel_id=channel("t.ret"))
        req.task_desc.slot_assignments.add(slot_name="/t/exc", channel_id=channel("t.exc"))
        op = wapi.execute_and_wait(req, timeout=60.0)
        resp = M.ExecuteResponse()
        assert op.done and op.response.Unpack(resp) and (resp.rc, resp.description) == (0, "Success")
        assert os.path.getsize(f"{sockdir}/t.ret") > 0
        wapi.close()
        priv.close()
        w.send_signal(signal.SIGTERM)
        assert w.wait(timeout=30) == 0, w.stderr.read()[-2000:]
        assert not os.path.exists(f"{sockdir}/w0.sock")
        cm.send_signal(signal.SIGTERM)
        assert cm.wait(timeout=30) == 0
        assert not os.path.exists(f"{sockdir}/cm.sock")
    finally:
        for p in (w, cm):
            if p is not None and p.poll() is None:
                p.kill()
