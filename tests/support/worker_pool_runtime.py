"""TEST INFRASTRUCTURE: a `Runtime` that executes a workflow's calls the way the reference's *remote* stack does —
pylzy's RemoteRuntime builds one task per call (slots "/{call.id}/{entry.id}", a command line carrying the pickled
ProcessingRequest, pylzy/lzy/api/v1/remote/runtime.py:303-420), lzy-service creates one channel per data URI
(CreateChannels.java:75-104), the graph executor releases tasks whose producers have finished
(BfsGraphProcessor.java:176-258) and each task goes to a worker as WorkerApi.Execute — composed from this
repository's own parts: `GraphDispatcher`, `ChannelManager`, `lzy_b200.worker.Worker`, `api.v1.startup`, the C++
slots runtime.  Slots run over the in-memory test backend and payloads are serializer byte streams, so the whole
stack runs without a GPU; the reference's workflow scenarios can then be asserted on the driver's CPU box too
(tests/test_workflow_host.py).  The product's own runtime for a GPU box is `B200Runtime`.
"""
from __future__ import annotations

import hashlib
import time
import uuid
from io import BytesIO
from typing import Any, Callable, Dict, List, Optional, Type

from lzy_b200 import _host as H
from lzy_b200.api.v1.exceptions import LzyExecutionException
from lzy_b200.api.v1.runtime import ProgressStep, Runtime
from lzy_b200.api.v1.snapshot import Snapshot, SnapshotEntry
from lzy_b200.api.v1.startup import ProcessingRequest
from lzy_b200.api.v1.utils.pickle import pickle
from lzy_b200.core.dataflow import plan_dataflow
from lzy_b200.proxy_result import Absence, Result
from lzy_b200.utils import tbpickle
from lzy_b200.worker import DIR_INPUT, DIR_OUTPUT, SlotSpec, TaskSpec, Worker
from tests.support import hostmem

tbpickle.install()


class BlobStorage:
    """the C++ host blob store of this process (mem:// URIs)"""

    def exists(self, uri: str) -> bool:
        return H.blob_exists(uri)

    def get(self, uri: str) -> bytes:
        return H.blob_get(uri)

    def put(self, uri: str, raw: bytes) -> None:
        H.blob_put(uri, raw)

    def delete(self, uri: str) -> None:
        H.blob_delete(uri)


class FileStorage:
    """file:// URIs — what every process of the box can reach.  Objects written by a worker carry the 128-byte
    frame header of the spill format; objects written by the client are plain byte streams."""

    @staticmethod
    def _path(uri: str) -> str:
        assert uri.startswith("file://"), uri
        return uri[7:]

    def exists(self, uri: str) -> bool:
        import os
        return os.path.exists(self._path(uri))

    def get(self, uri: str) -> bytes:
        with open(self._path(uri), "rb") as f:
            raw = f.read()
        return raw[128:] if raw[:8] == b"LZYB200\0" else raw

    def put(self, uri: str, raw: bytes) -> None:
        import os
        path = self._path(uri)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = f"{path}.tmp.{os.getpid()}"
        with open(tmp, "wb") as f:
            f.write(raw)
        os.replace(tmp, path)

    def delete(self, uri: str) -> None:
        import os
        try:
            os.unlink(self._path(uri))
        except FileNotFoundError:
            pass


class BlobSnapshot(Snapshot):
    """`DefaultSnapshot`'s rules (pylzy/lzy/api/v1/snapshot.py:93-191) over the C++ host blob store: the payload is
    the serializer's byte stream, the URI of local data is prefix + md5 of those bytes, equal data is stored once."""

    def __init__(self, registry, prefix: str, storage_name: str = "mem", storage=None):
        self.registry, self.prefix, self.storage_name = registry, prefix, storage_name
        self.storage = storage or BlobStorage()
        self.entries: Dict[str, SnapshotEntry] = {}
        self.store_counts: Dict[str, int] = {}

    def create_entry(self, name: str, typ: Type) -> SnapshotEntry:
        ser = self.registry.find_serializer_by_type(typ)
        if ser is None:
            raise TypeError(f"Cannot find serializer for type {typ}")
        if not ser.available():
            raise TypeError(f"Serializer for type {typ} is not available, please install {ser.requirements()}")
        e = SnapshotEntry(str(uuid.uuid4()), name, typ, ser.schema(typ), self.storage_name)
        self.entries[e.id] = e
        return e

    def get(self, entry_id: str) -> SnapshotEntry:
        return self.entries[entry_id]

    async def put_data(self, entry_id: str, data: Any) -> None:
        e = self.entries.get(entry_id)
        if e is None:
            raise ValueError(f"Entry with id={entry_id} does not exist")
        buf = BytesIO()
        self.registry.find_serializer_by_type(e.typ).serialize(data, buf)
        raw = buf.getvalue()
        e.data_hash = hashlib.md5(raw).hexdigest()
        e.storage_uri = self.prefix + "/" + e.data_hash
        if not self.storage.exists(e.storage_uri):
            self.storage.put(e.storage_uri, raw)
            self.store_counts[e.storage_uri] = self.store_counts.get(e.storage_uri, 0) + 1

    async def get_data(self, entry_id: str):
        e = self.entries.get(entry_id)
        if e is None:
            raise ValueError(f"Entry with id={entry_id} does not exist")
        try:
            uri = e.storage_uri
        except ValueError as err:
            return Absence(err)
        if not self.storage.exists(uri):
            return Absence(ValueError(f"Cannot find blob in storage by uri={uri}"))
        ser = self.registry.find_serializer_by_type(e.typ)
        return Result(ser.deserialize(BytesIO(self.storage.get(uri)), e.typ if isinstance(e.typ, type) else None))

    async def copy_data(self, from_entry_id: str, to_uri: str) -> None:
        self.storage.put(to_uri, self.storage.get(self.entries[from_entry_id].storage_uri))

    def uris(self) -> List[str]:
        return [e._storage_uri for e in self.entries.values() if e._storage_uri]


class WorkerPoolRuntime(Runtime):
    _live = 0  # runtimes that installed the in-memory backend: the last one to close removes it

    def __init__(self, n_workers: int = 2, user_id: str = "test-user", over_grpc: bool = False):
        """over_grpc: every worker sits behind its own WorkerApi / LongRunningService / LzySlotsApi server and is
        driven through the wire messages, as the reference's graph executor drives a worker VM"""
        self.n_workers, self.user_id, self.over_grpc = n_workers, user_id, over_grpc
        self._servers: List[Any] = []
        self._clients: Dict[int, Any] = {}
        self.storage = BlobStorage()
        self.cm: Optional[H.ChannelManager] = None
        self.workers: List[Worker] = []
        self.workflow = None
        self.execution_id: Optional[str] = None
        self.calls: List[Any] = []          # every call ever handed to exec (what RuntimeMock records)
        self.executed: List[str] = []       # callable names in the order tasks were released
        self.skipped_cached = 0
        self.dropped_duplicates = 0
        self._snapshots: List[BlobSnapshot] = []
        self._rr = 0
        self._owner = None
        self._generation = 0

    # ---- bring-up / teardown ---------------------------------------------------------------------------------
    def _ensure(self) -> None:
        if self.cm is not None:
            return
        hostmem.install()
        WorkerPoolRuntime._live += 1
        self.cm = H.ChannelManager(seed=1)
        H.pump_start(self.cm)

    def _allocate(self, workflow_name: str) -> None:
        """the allocator's part: a worker serves one (user, workflow) owner (WorkerApiImpl.java:232-283), so another
        workflow gets fresh workers"""
        if self.workers and self._owner == (self.user_id, workflow_name):
            return
        self._release_workers()
        self._generation += 1
        self.workers = [Worker(self.cm, 0, f"lzyb://pool-worker-{self._generation}-{i}", mount_factory=hostmem.HostmemMount)
                        for i in range(self.n_workers)]
        self._owner = (self.user_id, workflow_name)
        if self.over_grpc:
            from lzy_b200.grpc_api import LongRunningClient, WorkerApiClient, messages, serve
            for i, w in enumerate(self.workers):
                server, port = serve("127.0.0.1:0", worker=w)
                self._servers.append(server)
                self._clients[i] = (WorkerApiClient(f"127.0.0.1:{port}"), LongRunningClient(f"127.0.0.1:{port}"))
                self._clients[i][0].call("Init", messages.InitRequest(user_id=self.user_id, workflow_name=workflow_name))
        else:
            for w in self.workers:
                w.init(self.user_id, workflow_name)

    def _release_workers(self) -> None:
        for wapi, ops in self._clients.values():
            wapi.close()
            ops.close()
        self._clients.clear()
        for server in self._servers:
            server.stop(0)
        self._servers.clear()
        for w in self.workers:
            w.close()
        self.workers = []

    # ---- the private channel-manager API: a local matcher here, RPCs in the process-per-worker subclass ---------------
    def _get_or_create(self, workflow_name: str, execution_id: str, **peer) -> str:
        return self.cm.get_or_create(self.user_id, workflow_name, execution_id, **peer)

    def _destroy_all(self, execution_id: str, reason: str) -> None:
        self.cm.destroy_all(execution_id, reason)

    # ---- one task to one worker: direct calls, or the wire ------------------------------------------------------
    def _submit(self, index: int, task: TaskSpec) -> str:
        if not self.over_grpc:
            return self.workers[index].execute(task).id
        from lzy_b200.grpc_api import messages
        req = messages.ExecuteRequest(task_id=task.task_id, execution_id=task.execution_id, workflow_name=task.workflow_name,
                                      user_id=task.user_id)
        req.task_desc.operation.command = task.command
        req.task_desc.operation.name = task.name
        for k, v in (task.env or {}).items():
            req.task_desc.operation.env.env[k] = v
        for slot in task.slots:
            req.task_desc.operation.slots.add(name=slot.name, direction=slot.direction)
        for name, chan in task.slot_assignments.items():
            req.task_desc.slot_assignments.add(slot_name=name, channel_id=chan)
        return self._clients[index][0].call("Execute", req).id

    def _poll(self, index: int, op_id: str):
        """None while the task runs, else (ok, description)"""
        if not self.over_grpc:
            op = self.workers[index].get_operation(op_id)
            if not op.done:
                return None
            ok = op.error is None and op.response is not None and op.response.rc == 0
            return ok, (op.response.description if op.response else str(op.error))
        from lzy_b200.grpc_api import messages
        from lzy_b200.grpc_api.protos import GetOperationRequest
        op = self._clients[index][1].call("Get", GetOperationRequest(operation_id=op_id))
        if not op.done:
            return None
        if op.WhichOneof("result") != "response":
            return False, op.error.message
        resp = messages.ExecuteResponse()
        op.response.Unpack(resp)
        return resp.rc == 0, resp.description

    def close(self) -> None:
        if self.cm is None:
            return
        self._release_workers()
        H.pump_stop(self.cm)
        self.cm.close()
        self.cm = None
        for snap in self._snapshots:
            for uri in snap.uris():
                self.storage.delete(uri)
        self._snapshots.clear()
        WorkerPoolRuntime._live -= 1
        if WorkerPoolRuntime._live == 0:
            hostmem.uninstall()

    # ---- Runtime ABC ---------------------------------------------------------------------------------------------
    async def storage(self):
        return None

    def snapshot(self, workflow, storage_uri_prefix: str) -> Snapshot:
        snap = BlobSnapshot(workflow.owner.serializer_registry, storage_uri_prefix, storage=self.storage)
        self._snapshots.append(snap)
        return snap

    async def start(self, workflow) -> str:
        self._ensure()
        self._allocate(workflow.name)
        self.workflow = workflow
        self.execution_id = str(uuid.uuid4())
        return self.execution_id

    async def abort(self) -> None:
        self._finish("aborted")

    async def finish(self) -> None:
        self._finish("finished")

    def _finish(self, reason: str) -> None:
        if self.cm is not None and self.execution_id:
            for w in self.workers:
                w.release_execution(self.execution_id)
            self._destroy_all(self.execution_id, reason)
            self.execution_id = None

    async def exec(self, calls, progress: Callable[[ProgressStep], None]) -> None:
        import asyncio
        self.calls.extend(calls)
        progress(ProgressStep.WAITING)
        try:
            await asyncio.get_running_loop().run_in_executor(None, self._exec_graph, list(calls), progress)
        except BaseException:
            progress(ProgressStep.FAILED)
            raise
        progress(ProgressStep.COMPLETED)

    # ---- the graph ---------------------------------------------------------------------------------------------------
    def _exec_graph(self, calls: List[Any], progress) -> None:
        wf, cm = self.workflow, self.cm
        snap: BlobSnapshot = wf.snapshot
        registry = wf.owner.serializer_registry
        exec_id = self.execution_id

        # ops whose outputs are all already supplied are dropped (lzy-service DataFlowGraph.java:51-72)
        plan = plan_dataflow(calls, lambda eid: snap.get(eid).storage_uri, self.storage.exists)
        self.skipped_cached += len(plan.cached)
        self.dropped_duplicates += len(plan.duplicates)
        todo = plan.calls
        if not todo:
            return
        producer_of = plan.supplier

        # CreateChannels: one channel per data URI; results get a storage consumer, existing data a storage producer
        channel: Dict[str, str] = {}
        for c in todo:
            for eid in list(c.entry_ids) + [c.exception_id]:
                channel[eid] = self._get_or_create(wf.name, exec_id, consumer_uri=snap.get(eid).storage_uri)
        for eid in plan.supplier:  # results of dropped duplicates: the channel of the URI they share with their supplier
            if eid not in channel:
                channel[eid] = self._get_or_create(wf.name, exec_id, consumer_uri=snap.get(eid).storage_uri)
        tasks: List[TaskSpec] = []
        edges = set()
        for i, c in enumerate(todo):
            slots, assign = [], {}
            args_paths, kwargs_paths, outs = [], {}, []

            def input_slot(eid):
                if eid not in channel:
                    uri = snap.get(eid).storage_uri
                    if not self.storage.exists(uri):
                        raise ValueError(f"Cannot find blob in storage by uri={uri}")
                    channel[eid] = self._get_or_create(wf.name, exec_id, producer_uri=uri)
                if eid in producer_of and producer_of[eid] != i:
                    edges.add((producer_of[eid], i))
                path = f"/{c.id}/{eid}"
                if path not in assign:
                    assign[path] = channel[eid]
                    slots.append(SlotSpec(path, DIR_INPUT))
                return snap.get(eid).typ, path

            for eid in c.arg_entry_ids:
                args_paths.append(input_slot(eid))
            for name, eid in c.kwarg_entry_ids.items():
                kwargs_paths[name] = input_slot(eid)
            for eid in c.entry_ids:
                path = f"/{c.id}/{eid}"
                assign[path] = channel[eid]
                slots.append(SlotSpec(path, DIR_OUTPUT))
                outs.append((snap.get(eid).typ, path))
            exc_path = f"/{c.id}/{c.exception_id}"
            assign[exc_path] = channel[c.exception_id]
            slots.append(SlotSpec(exc_path, DIR_OUTPUT))
            request = ProcessingRequest({}, registry.imports(), c.signature.func.callable, args_paths, kwargs_paths, outs,
                                        (tuple, exc_path), lazy_arguments=c.lazy_arguments)
            command = " ".join(["python -u", "$(python -c 'import lzy.api.v1.startup as startup; print(startup.__file__)')",
                                pickle(request)])
            tasks.append(TaskSpec(task_id=c.id, execution_id=exec_id, workflow_name=wf.name, user_id=self.user_id,
                                  command=command, slots=slots, slot_assignments=assign, name=c.callable_name,
                                  env=c.get_env_vars()))

        graph = H.GraphDispatcher(len(todo), sorted(edges), allow_cycles=False)
        running: Dict[int, Any] = {}
        failed: Optional[int] = None
        progress(ProgressStep.EXECUTING)
        try:
            deadline = time.time() + 120
            while graph.status not in (H.G_COMPLETED, H.G_FAILED):
                if time.time() > deadline:
                    raise LzyExecutionException("graph did not finish in time")
                for t in graph.next():
                    index = self._rr % len(self.workers)
                    self._rr += 1
                    self.executed.append(todo[t].callable_name)
                    running[t] = (index, self._submit(index, tasks[t]))
                for t, (index, op_id) in list(running.items()):
                    outcome = self._poll(index, op_id)
                    if outcome is None:
                        continue
                    del running[t]
                    ok, description = outcome
                    if not ok and failed is None:
                        failed = t
                    graph.task_done(t, ok, "" if ok else description)
                time.sleep(0.001)
            if failed is not None or graph.status == H.G_FAILED:
                self._raise_failure(todo[failed] if failed is not None else None, snap, graph)
        finally:
            deadline = time.time() + 30
            for index, op_id in running.values():  # tasks still in flight when the graph stopped
                while self._poll(index, op_id) is None and time.time() < deadline:
                    time.sleep(0.002)
            graph.close()

    def _raise_failure(self, call, snap: BlobSnapshot, graph) -> None:
        """RemoteRuntime.exec's Failed branch (remote/runtime.py:190-205): re-raise the op's own exception"""
        if call is not None:
            uri = snap.get(call.exception_id).storage_uri
            deadline = time.time() + 10
            while not self.storage.exists(uri) and time.time() < deadline:  # the exception slot uploads asynchronously
                time.sleep(0.005)
            if self.storage.exists(uri):
                ser = snap.registry.find_serializer_by_type(tuple)
                info = ser.deserialize(BytesIO(self.storage.get(uri)), tuple)
                if info is not None:
                    raise info[1].with_traceback(info[2])
        _, description, _ = graph.error()
        raise LzyExecutionException(f"Failed executing graph: {description}")


class ProcessPoolRuntime(WorkerPoolRuntime):
    """The same runtime with every part in its own process, as the reference deploys it: a channel-manager service
    (`python -m lzy_b200.grpc_api channel-manager`), N worker processes (tests/support/remote_worker_api.py: WorkerApi +
    LzySlotsApi + LzybPeerApi, in-memory backend with its shared-memory export), this process as the client that
    builds tasks and polls operations.  Data lives under file:// URIs so that every process reaches it: the client
    writes arguments there, workers upload results there and pull each other's outputs slot to slot."""

    def __init__(self, root_dir: str, n_workers: int = 2, user_id: str = "test-user"):
        super().__init__(n_workers=n_workers, user_id=user_id, over_grpc=True)
        self.storage = FileStorage()
        self.root_dir = root_dir
        self._cm_proc = None
        self._worker_procs: List[Any] = []
        self._priv = None

    def _ensure(self) -> None:
        if self._cm_proc is not None:
            return
        import subprocess
        import sys
        from pathlib import Path

        from lzy_b200.grpc_api import ChannelManagerPrivateClient
        self._repo = Path(__file__).resolve().parent.parent.parent
        self._cm_proc = subprocess.Popen([sys.executable, "-m", "lzy_b200.grpc_api", "channel-manager", "--port", "0", "--seed", "1"],
                                         cwd=str(self._repo), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        self._cm_target = self._cm_proc.stdout.readline().strip().rsplit(" ", 1)[-1]
        self._priv = ChannelManagerPrivateClient(self._cm_target)
        self.cm = self._priv  # (only its presence is checked by the base class)

    def _allocate(self, workflow_name: str) -> None:
        if self._worker_procs and self._owner == (self.user_id, workflow_name):
            return
        import socket
        import subprocess
        import sys

        from lzy_b200.grpc_api import LongRunningClient, WorkerApiClient, messages
        self._release_workers()
        self._owner = (self.user_id, workflow_name)
        for i in range(self.n_workers):
            with socket.socket() as sock:
                sock.bind(("127.0.0.1", 0))
                port = sock.getsockname()[1]
            proc = subprocess.Popen([sys.executable, str(self._repo / "tests/support/remote_worker_api.py"), self._cm_target, str(port)],
                                    cwd=str(self._repo), stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            assert proc.stdout.readline().startswith("READY")
            self._worker_procs.append(proc)
            self._clients[i] = (WorkerApiClient(f"127.0.0.1:{port}"), LongRunningClient(f"127.0.0.1:{port}"))
            self._clients[i][0].call("Init", messages.InitRequest(user_id=self.user_id, workflow_name=workflow_name))
        self.workers = list(range(self.n_workers))  # placeholders: the base class only counts them

    def _release_workers(self) -> None:
        for wapi, ops in self._clients.values():
            wapi.close()
            ops.close()
        self._clients.clear()
        for proc in self._worker_procs:
            try:
                proc.stdin.write("bye\n")
                proc.stdin.flush()
            except Exception:  # noqa: BLE001
                pass
        for proc in self._worker_procs:
            try:
                proc.wait(timeout=20)
            except Exception:  # noqa: BLE001
                proc.kill()
        self._worker_procs = []
        self.workers = []

    def _get_or_create(self, workflow_name: str, execution_id: str, producer_uri=None, consumer_uri=None) -> str:
        from lzy_b200.grpc_api import messages
        req = messages.GetOrCreateRequest(execution_id=execution_id, workflow_name=workflow_name, user_id=self.user_id)
        if producer_uri is not None:
            req.producer.storage_uri = producer_uri
        else:
            req.consumer.storage_uri = consumer_uri
        return self._priv.call("GetOrCreate", req).channel_id

    def _destroy_all(self, execution_id: str, reason: str) -> None:
        from lzy_b200.grpc_api import messages
        self._priv.call("DestroyAll", messages.DestroyAllRequest(execution_id=execution_id, reason=reason))

    def _finish(self, reason: str) -> None:
        if self._priv is not None and self.execution_id:
            self._destroy_all(self.execution_id, reason)
            self.execution_id = None

    def close(self) -> None:
        if self._cm_proc is None:
            return
        self._release_workers()
        self._priv.close()
        self._cm_proc.terminate()
        self._cm_proc.wait(timeout=15)
        self._cm_proc, self.cm = None, None
        import shutil
        shutil.rmtree(self.root_dir, ignore_errors=True)
