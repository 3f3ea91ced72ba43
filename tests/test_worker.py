"""`lzy_b200.worker.Worker` — WorkerApi.Init / Execute + LongRunningService.Get — against what
`WorkerApiImpl` does (lzy/worker/src/main/java/ai/lzy/worker/WorkerApiImpl.java:85-300) and what the reference's
worker test pins about availability (lzy/worker/src/test/java/ai/lzy/worker/WorkerTests.java:253-330: the slots
API exists only after Init, unknown operation → NOT_FOUND, Execute without ownership → PERMISSION_DENIED; the IAM
half of that test is out of scope).  Staging order and return codes: SURVEY Appendix A.4.  Slots run over the
in-memory test backend; the op body goes through `startup.main` with a mount on that backend."""
import time
from io import BytesIO

import grpc
import pytest

from lzy_b200 import _host as H
from lzy_b200.api.v1.startup import ProcessingRequest
from lzy_b200.api.v1.utils.pickle import pickle
from lzy_b200.grpc_api import LongRunningClient, SlotsApiClient, WorkerApiClient, messages as M, serve
from lzy_b200.grpc_api.protos import GetOperationRequest
from lzy_b200.serialization.registry import LzySerializerRegistry
from lzy_b200.worker import (DIR_INPUT, DIR_OUTPUT, InstallationError, ReturnCodes, SlotSpec, TaskSpec, Worker,
                             WorkerStatusError)
from tests.support import hostmem

SER = LzySerializerRegistry()


def enc(typ, value) -> bytes:
    buf = BytesIO()
    SER.find_serializer_by_type(typ).serialize(value, buf)
    return buf.getvalue()


def dec(typ, raw: bytes):
    return SER.find_serializer_by_type(typ).deserialize(BytesIO(raw), typ if isinstance(typ, type) else None)


@pytest.fixture
def env():
    hostmem.install()
    cm = H.ChannelManager(seed=5)
    H.pump_start(cm)
    worker = Worker(cm, 0, "lzyb://worker-0", mount_factory=hostmem.HostmemMount)
    yield cm, worker
    worker.close()
    H.pump_stop(cm)
    cm.close()
    hostmem.uninstall()
    for uri in list(_blobs):
        H.blob_delete(uri)
    _blobs.clear()
    assert hostmem.load().hostmem_live_slots() == 0, "a slot frame leaked"


_blobs = set()


def make_task(cm, op, args, execution="exec-1", task_id="t1", n_out=1, out_types=(str,), user="uid", wf="wf"):
    """what lzy-service's CreateChannels/BuildTasks + pylzy's __build_graph produce for one call: a channel per
    entry (storage producer for existing data, storage consumer for results), slots "/{call}/{entry}", the
    startup command line with the pickled request as its last token (remote/runtime.py:326-381)"""
    slots, assign, args_paths = [], {}, []
    for i, (typ, value) in enumerate(args):
        uri = f"mem://{execution}/{task_id}/arg{i}"
        H.blob_put(uri, enc(typ, value))
        _blobs.add(uri)
        path = f"/{task_id}/arg{i}"
        assign[path] = cm.get_or_create(user, wf, execution, producer_uri=uri)
        slots.append(SlotSpec(path, DIR_INPUT))
        args_paths.append((typ, path))
    outs = []
    for i in range(n_out):
        path, uri = f"/{task_id}/ret{i}", f"mem://{execution}/{task_id}/ret{i}"
        _blobs.add(uri)
        assign[path] = cm.get_or_create(user, wf, execution, consumer_uri=uri)
        slots.append(SlotSpec(path, DIR_OUTPUT))
        outs.append((out_types[i], path))
    exc_path, exc_uri = f"/{task_id}/exc", f"mem://{execution}/{task_id}/exc"
    _blobs.add(exc_uri)
    assign[exc_path] = cm.get_or_create(user, wf, execution, consumer_uri=exc_uri)
    slots.append(SlotSpec(exc_path, DIR_OUTPUT))
    req = ProcessingRequest({}, SER.imports(), op, args_paths, {}, outs, (tuple, exc_path), lazy_arguments=False)
    command = " ".join(["python -u", "$(python -c 'import lzy.api.v1.startup as startup; print(startup.__file__)')", pickle(req)])
    return TaskSpec(task_id=task_id, execution_id=execution, workflow_name=wf, user_id=user, command=command, slots=slots,
                    slot_assignments=assign)


def wait_blob(uri, timeout=5.0) -> bytes:
    deadline = time.time() + timeout
    while not H.blob_exists(uri):
        assert time.time() < deadline, f"{uri} never uploaded"
        time.sleep(0.005)
    return H.blob_get(uri)


def concat(a: str, b: str) -> str:
    return a + b


def boom(a: str) -> str:
    raise ValueError("op failed: " + a)


def test_execute_needs_init_and_the_owner(env):
    cm, worker = env
    task = make_task(cm, concat, [(str, "4"), (str, "2")])
    assert worker.slots is None  # the slots server exists only after Init
    with pytest.raises(WorkerStatusError, match="PERMISSION_DENIED"):
        worker.execute(task)
    worker.init("uid", "wf")
    assert worker.slots is not None
    with pytest.raises(WorkerStatusError, match="PERMISSION_DENIED"):
        worker.init("hacker-user", "wf")
    with pytest.raises(WorkerStatusError, match="PERMISSION_DENIED"):
        worker.execute(make_task(cm, concat, [(str, "a"), (str, "b")], task_id="t2", user="hacker-user"))
    worker.init("uid", "wf")  # the same owner may re-init
    with pytest.raises(WorkerStatusError, match="NOT_FOUND"):
        worker.get_operation("1")


def test_execute_stages_inputs_runs_the_op_and_uploads_outputs(env):
    cm, worker = env
    worker.init("uid", "wf")
    task = make_task(cm, concat, [(str, "4"), (str, "2")])
    op = worker.execute(task, idempotency_key="k1")
    assert op.created_by == "t1" and op.description == "Worker, executionId: exec-1, taskId: t1"
    assert worker.execute(task, idempotency_key="k1").id == op.id  # replay of the same request: same operation
    done = worker.wait_operation(op.id)
    assert done.done and done.error is None
    assert (done.response.rc, done.response.description) == (ReturnCodes.SUCCESS, "Success")
    # afterExecution returned ⇒ every output is bound and uploaded to its storage consumer
    assert dec(str, H.blob_get("mem://exec-1/t1/ret0")) == "42"
    assert dec(type(None), H.blob_get("mem://exec-1/t1/exc")) is None
    # and a downstream task on the same worker gets it slot → slot
    t2 = make_task(cm, concat, [(str, "<"), (str, ">")], task_id="t2")
    path = "/t2/arg0"
    t2.slot_assignments[path] = task.slot_assignments["/t1/ret0"]  # consume t1's result channel instead
    done2 = worker.wait_operation(worker.execute(t2).id)
    assert done2.response.rc == 0
    assert dec(str, H.blob_get("mem://exec-1/t2/ret0")) == "42>"
    worker.release_execution("exec-1")


def test_failing_op_is_rc_1_and_its_exception_slot_still_reaches_storage(env):
    cm, worker = env
    worker.init("uid", "wf")
    done = worker.wait_operation(worker.execute(make_task(cm, boom, [(str, "x")])).id)
    assert done.response.rc == 1
    assert done.response.description == "Error while executing command on worker. See your stdout/stderr to see more info."
    exc = dec(tuple, wait_blob("mem://exec-1/t1/exc"))
    assert exc[0] is ValueError and exc[1].args == ("op failed: x",)
    assert not H.blob_exists("mem://exec-1/t1/ret0")


def test_missing_slot_assignment_is_internal_error_98(env):
    cm, worker = env
    worker.init("uid", "wf")
    task = make_task(cm, concat, [(str, "a"), (str, "b")])
    del task.slot_assignments["/t1/ret0"]
    done = worker.wait_operation(worker.execute(task).id)
    assert (done.response.rc, done.response.description) == (ReturnCodes.INTERNAL_ERROR, "Internal error")


def test_environment_installation_error_is_97(env):
    cm, worker = env

    def no_conda(spec):
        raise InstallationError("conda env create failed")

    worker.env_factory = no_conda
    worker.init("uid", "wf")
    done = worker.wait_operation(worker.execute(make_task(cm, concat, [(str, "a"), (str, "b")])).id)
    assert (done.response.rc, done.response.description) == (ReturnCodes.ENVIRONMENT_INSTALLATION_ERROR, "conda env create failed")
    worker.env_factory = lambda spec: 1 / 0
    done = worker.wait_operation(worker.execute(make_task(cm, concat, [(str, "a"), (str, "b")], task_id="t2")).id)
    assert (done.response.rc, done.response.description) == (ReturnCodes.INTERNAL_ERROR, "Internal error")


def test_input_that_cannot_be_delivered_is_internal_error(env):
    cm, worker = env
    worker.init("uid", "wf")
    task = make_task(cm, concat, [(str, "a"), (str, "b")])
    H.blob_delete("mem://exec-1/t1/arg1")  # the only producer of that channel can never deliver
    done = worker.wait_operation(worker.execute(task).id)
    assert (done.response.rc, done.response.description) == (ReturnCodes.INTERNAL_ERROR, "Internal error")
    assert cm.aborted("exec-1").startswith("No more producers")


def test_worker_api_over_grpc(env):
    """WorkerTests.apiAvailability without IAM, then one Execute polled through LongRunningService.Get."""
    cm, worker = env
    server, port = serve("127.0.0.1:0", worker=worker)
    target = f"127.0.0.1:{port}"
    wapi, ops, sapi = WorkerApiClient(target), LongRunningClient(target), SlotsApiClient(target)
    try:
        def code(fn):
            with pytest.raises(grpc.RpcError) as e:
                r = fn()
                if hasattr(r, "__next__"):
                    next(r)
            return e.value.code()

        # before Init: no slots server, no owner
        assert code(lambda: sapi.call("StartTransfer", M.StartTransferRequest())) == grpc.StatusCode.UNAVAILABLE
        assert code(lambda: sapi.call("Read", M.ReadDataRequest())) == grpc.StatusCode.UNAVAILABLE
        assert code(lambda: wapi.call("Execute", M.ExecuteRequest())) == grpc.StatusCode.PERMISSION_DENIED
        wapi.call("Init", M.InitRequest(user_id="uid", workflow_name="wf", worker_subject_name="w", worker_private_key="k"))
        assert code(lambda: ops.call("Get", GetOperationRequest(operation_id="1"))) == grpc.StatusCode.NOT_FOUND
        assert code(lambda: sapi.call("StartTransfer", M.StartTransferRequest())) == grpc.StatusCode.NOT_FOUND
        assert code(lambda: sapi.call("Read", M.ReadDataRequest())) == grpc.StatusCode.NOT_FOUND
        assert code(lambda: wapi.call("Init", M.InitRequest(user_id="hacker-user", workflow_name="wf"))) == grpc.StatusCode.PERMISSION_DENIED

        task = make_task(cm, concat, [(str, "4"), (str, "2")])
        req = M.ExecuteRequest(task_id=task.task_id, execution_id=task.execution_id, workflow_name="wf", user_id="uid")
        req.task_desc.operation.command = task.command
        req.task_desc.operation.name = "concat"
        for s in task.slots:
            req.task_desc.operation.slots.add(name=s.name, direction=s.direction)
        for name, chan in task.slot_assignments.items():
            req.task_desc.slot_assignments.add(slot_name=name, channel_id=chan)
        op = wapi.call("Execute", req, idempotency_key="idk-1")
        assert op.created_by == "t1" and wapi.call("Execute", req, idempotency_key="idk-1").id == op.id
        deadline = time.time() + 30
        while not op.done:
            assert time.time() < deadline
            time.sleep(0.01)
            op = ops.call("Get", GetOperationRequest(operation_id=op.id))
        assert op.WhichOneof("result") == "response"
        resp = M.ExecuteResponse()
        assert op.response.Unpack(resp) and (resp.rc, resp.description) == (0, "Success")
        assert dec(str, H.blob_get("mem://exec-1/t1/ret0")) == "42"
        # a legacy consumer can Read the worker's output slot over the same port (2 MiB chunks + EOS)
        assert dec(str, sapi.read_all("/t1/ret0")) == "42"
    finally:
        for c in (wapi, ops, sapi):
            c.close()
        server.stop(0)


def _slow(a: str) -> str:
    time.sleep(1.0)
    return a + "!"


def test_get_with_the_wait_header_is_answered_when_the_task_completes(env):
    """`lzyb-wait-ms` on LongRunningService.Get (grpc_api/server.py): without it the answer comes at once, done or not, as from
    the reference's operation service; with it a running operation is answered when it completes — one call, no polling —
    and an expired wait returns the operation as it stands; nonsense in the header is INVALID_ARGUMENT."""
    cm, worker = env
    server, port = serve("127.0.0.1:0", worker=worker)
    wapi, ops = WorkerApiClient(f"127.0.0.1:{port}"), LongRunningClient(f"127.0.0.1:{port}")
    try:
        wapi.call("Init", M.InitRequest(user_id="uid", workflow_name="wf"))
        task = make_task(cm, _slow, [(str, "x")])
        req = M.ExecuteRequest(task_id=task.task_id, execution_id=task.execution_id, workflow_name="wf", user_id="uid")
        req.task_desc.operation.command = task.command
        for s in task.slots:
            req.task_desc.operation.slots.add(name=s.name, direction=s.direction)
        for name, chan in task.slot_assignments.items():
            req.task_desc.slot_assignments.add(slot_name=name, channel_id=chan)
        t0 = time.monotonic()
        op = wapi.call("Execute", req)
        get = GetOperationRequest(operation_id=op.id)
        assert not ops.call("Get", get).done and time.monotonic() - t0 < 0.8          # a stock Get never waits
        assert not ops.call("Get", get, metadata=(("lzyb-wait-ms", "50"),)).done       # a wait that expires: as it stands
        with pytest.raises(grpc.RpcError) as e:
            ops.call("Get", get, metadata=(("lzyb-wait-ms", "soon"),))
        assert e.value.code() == grpc.StatusCode.INVALID_ARGUMENT
        done = ops.wait(op.id, timeout=30.0)                                            # one call, answered on completion
        took = time.monotonic() - t0
        assert done.done and 0.9 < took < 10.0
        resp = M.ExecuteResponse()
        assert done.response.Unpack(resp) and (resp.rc, resp.description) == (0, "Success")
        assert dec(str, H.blob_get("mem://exec-1/t1/ret0")) == "x!"
        assert ops.wait(op.id).done                                                     # already done: at once
        # Execute itself honours the header: the operation comes back completed, no Get at all
        task2 = make_task(cm, _slow, [(str, "y")], task_id="t2")
        req2 = M.ExecuteRequest(task_id=task2.task_id, execution_id=task2.execution_id, workflow_name="wf", user_id="uid")
        req2.task_desc.operation.command = task2.command
        for s in task2.slots:
            req2.task_desc.operation.slots.add(name=s.name, direction=s.direction)
        for name, chan in task2.slot_assignments.items():
            req2.task_desc.slot_assignments.add(slot_name=name, channel_id=chan)
        t1 = time.monotonic()
        done2 = wapi.execute_and_wait(req2, timeout=30.0)
        assert done2.done and time.monotonic() - t1 > 0.9 and done2.created_by == "t2"
        resp2 = M.ExecuteResponse()
        assert done2.response.Unpack(resp2) and resp2.rc == 0 and dec(str, H.blob_get("mem://exec-1/t2/ret0")) == "y!"
        with pytest.raises(grpc.RpcError) as e:
            ops.call("Get", GetOperationRequest(operation_id="nope"), metadata=(("lzyb-wait-ms", "50"),))
        assert e.value.code() == grpc.StatusCode.NOT_FOUND
    finally:
        wapi.close()
        ops.close()
        server.stop(0)


def test_diamond_graph_driven_through_two_workers_over_grpc(env):
    """What the reference's graph executor does with its workers, end to end on this repository's pieces: channels
    from CreateChannels' rule (storage consumer per result, storage producer per existing input), tasks released by
    the dispatcher when their producers have succeeded (BfsGraphProcessor), each sent to a worker as
    WorkerApi.Execute and polled through LongRunningService.Get; b and c pull a's result slot to slot from the
    other worker, d pulls from both."""
    cm, w0 = env
    w1 = Worker(cm, 0, "lzyb://worker-1", mount_factory=hostmem.HostmemMount)
    servers, clients = [], []
    try:
        for w in (w0, w1):
            server, port = serve("127.0.0.1:0", worker=w)
            servers.append(server)
            clients.append((WorkerApiClient(f"127.0.0.1:{port}"), LongRunningClient(f"127.0.0.1:{port}")))
            clients[-1][0].call("Init", M.InitRequest(user_id="uid", workflow_name="wf"))

        # a: "x" -> "xa";  b, c: t -> t + "b" / "c";  d: (l, r) -> l + "|" + r
        def tag(suffix):
            def f(t: str) -> str:
                return t + suffix
            return f

        def join(l: str, r: str) -> str:
            return l + "|" + r

        tasks = [make_task(cm, tag("a"), [(str, "x")], task_id="a"),
                 make_task(cm, tag("b"), [(str, "?")], task_id="b"),
                 make_task(cm, tag("c"), [(str, "?")], task_id="c"),
                 make_task(cm, join, [(str, "?"), (str, "?")], task_id="d")]
        tasks[1].slot_assignments["/b/arg0"] = tasks[0].slot_assignments["/a/ret0"]
        tasks[2].slot_assignments["/c/arg0"] = tasks[0].slot_assignments["/a/ret0"]
        tasks[3].slot_assignments["/d/arg0"] = tasks[1].slot_assignments["/b/ret0"]
        tasks[3].slot_assignments["/d/arg1"] = tasks[2].slot_assignments["/c/ret0"]
        placement = [0, 1, 0, 1]
        graph = H.GraphDispatcher(4, [(0, 1), (0, 2), (1, 3), (2, 3)], allow_cycles=False)
        running, order = {}, []
        deadline = time.time() + 60
        while graph.status not in (H.G_COMPLETED, H.G_FAILED):
            assert time.time() < deadline
            for t in graph.next():
                task = tasks[t]
                req = M.ExecuteRequest(task_id=task.task_id, execution_id=task.execution_id, workflow_name="wf", user_id="uid")
                req.task_desc.operation.command = task.command
                for s in task.slots:
                    req.task_desc.operation.slots.add(name=s.name, direction=s.direction)
                for name, chan in task.slot_assignments.items():
                    req.task_desc.slot_assignments.add(slot_name=name, channel_id=chan)
                running[t] = clients[placement[t]][0].call("Execute", req).id
                order.append(t)
            for t, op_id in list(running.items()):
                op = clients[placement[t]][1].call("Get", GetOperationRequest(operation_id=op_id))
                if op.done:
                    resp = M.ExecuteResponse()
                    assert op.response.Unpack(resp)
                    graph.task_done(t, resp.rc == 0, resp.description)
                    del running[t]
            time.sleep(0.002)
        assert graph.status == H.G_COMPLETED
        assert order[0] == 0 and order[-1] == 3 and set(order[1:3]) == {1, 2}
        assert dec(str, H.blob_get("mem://exec-1/d/ret0")) == "xab|xac"
        graph.close()
    finally:
        for wapi, ops in clients:
            wapi.close()
            ops.close()
        for s in servers:
            s.stop(0)
        w1.close()
