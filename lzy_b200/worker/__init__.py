"""The worker's staging step either side of the slot data plane — ``WorkerApi.Init`` / ``WorkerApi.Execute``
(lzy/worker-api/src/main/proto/ai/lzy/v1/worker/worker-service.proto:14-50) as the reference's
``WorkerApiImpl`` performs them (lzy/worker/src/main/java/ai/lzy/worker/WorkerApiImpl.java:85-300), so the
reference's scheduler / graph executor can drive a per-GPU worker of this repository unchanged:

* ``Init`` binds the worker to one (user, workflow) owner and only then creates the slots server
  (WorkerApiImpl.java:232-283); another owner is PERMISSION_DENIED; re-``Init`` by the same owner waits for
  the active execution to finish;
* ``Execute`` returns a long-running operation at once and stages the task in the background
  (WorkerApiImpl.java:85-141): every declared slot needs a channel assignment (else rc 98), environment
  (rc 97 on an installation error), ``slots.context(...)`` → ``beforeExecution()`` (all inputs delivered)
  → run the command → on rc 0 ``afterExecution()`` (all outputs bound / uploaded) → ``ExecuteResponse{rc,
  description}`` with the reference's texts (WorkerApiImpl.java:143-227; codes
  lzy/model/src/main/java/ai/lzy/model/ReturnCodes.java:3-7).

What changes is the command: the reference spawns ``python startup.py <request>`` against a directory of slot
files; here the task's request (the last token of the same command line, pylzy/lzy/api/v1/remote/
runtime.py:375-381) runs in-process on the worker's GPU through ``lzy_b200.api.v1.startup.main`` with the
execution context as its mount, because the argument and result frames live in this process's HBM.  Conda /
docker environment installation, Kafka log shipping and IAM tokens are out of scope; the
environment step is a pluggable callable so its failure path (rc 97) is still exercised.
"""
from __future__ import annotations

import contextlib
import dataclasses
import logging
import os
import sys
import threading
import time
import traceback
import uuid
from typing import Any, Callable, Dict, List, Mapping, Optional, Sequence, Tuple

from .. import _host as H
from ..env import applied_env

_LOG = logging.getLogger(__name__)

DIR_INPUT, DIR_OUTPUT = H.DIR_INPUT, H.DIR_OUTPUT


class ReturnCodes:
    """lzy/model/src/main/java/ai/lzy/model/ReturnCodes.java:3-7"""
    SUCCESS = 0
    ENVIRONMENT_INSTALLATION_ERROR = 97
    INTERNAL_ERROR = 98


class InstallationError(Exception):
    """Environment.InstallationException — the env step could not prepare what the op asked for."""


class WorkerStatusError(Exception):
    """An RPC-level refusal (gRPC status name + description), as opposed to a task result."""

    def __init__(self, code: str, message: str = ""):
        super().__init__(f"{code}: {message}" if message else code)
        self.code, self.message = code, message


@dataclasses.dataclass
class SlotSpec:
    name: str
    direction: int  # DIR_INPUT / DIR_OUTPUT (LMS.Slot.Direction)


@dataclasses.dataclass
class TaskSpec:
    """LWS.ExecuteRequest flattened (worker-service.proto:40-47, operation.proto:12-35)."""
    task_id: str
    execution_id: str
    workflow_name: str
    user_id: str
    command: str
    slots: Sequence[SlotSpec]
    slot_assignments: Mapping[str, str]  # slot name → channel id
    env: Any = None
    name: str = ""


@dataclasses.dataclass
class ExecuteResult:
    rc: int
    description: str


@dataclasses.dataclass
class Operation:
    """ai.lzy.v1.longrunning.Operation as far as WorkerApi uses it (long-running/operation.proto:20-34)."""
    id: str
    description: str
    created_by: str
    created_at: float
    modified_at: float
    done: bool = False
    response: Optional[ExecuteResult] = None
    error: Optional[Tuple[int, str]] = None  # (grpc code number, message)

    def snapshot(self) -> "Operation":
        return dataclasses.replace(self)


MountFactory = Callable[[H.ExecutionContext], Any]
EnvFactory = Callable[[Any], Any]
CommandRunner = Callable[[str, H.ExecutionContext, "Worker"], int]


def _env_vars(spec: Any) -> Dict[str, str]:
    """a plain mapping, or an ``LME.EnvSpec`` message whose ``env`` field is the map (env.proto:10-21)"""
    if spec is None:
        return {}
    inner = getattr(spec, "env", None)
    return dict(inner) if inner is not None and not callable(inner) else dict(spec) if isinstance(spec, Mapping) else {}


def _phase_marker(task_id: str) -> Callable[[str], None]:
    """LZYB_TRACE=1: one stderr line per staging phase of a task (milliseconds since the previous mark) — where a task's
    wall time goes between Execute and its response; otherwise a no-op"""
    if not os.environ.get("LZYB_TRACE"):
        return lambda phase: None
    last = [time.perf_counter()]

    def mark(phase: str) -> None:
        now = time.perf_counter()
        print(f"LZYB_TRACE task={task_id} phase={phase!r} ms={1e3 * (now - last[0]):.3f}", file=sys.stderr, flush=True)
        last[0] = now

    return mark


def run_startup_in_process(command: str, ctx: H.ExecutionContext, worker: "Worker") -> int:
    """The command a pylzy client builds is ``python -u <startup.py> <pickled request>``
    (remote/runtime.py:375-381); the request is its last token.  Exit code as a Python process would report
    it: 0, or 1 when the entry point raised (the op's exception is already in its exception slot)."""
    from ..api.v1 import startup
    from ..utils import op_scope
    from ..utils.argv_codec import request_token
    token = request_token(command)
    with worker.device_scope():
        mount = worker.mount_factory(ctx)
        op_scope.enter()
        try:
            startup.main(token, mount)
            return 0
        except Exception:
            traceback.print_exc(file=sys.stderr)
            return 1
        finally:
            op_scope.leave()


class Worker:
    """One per GPU.  ``init`` / ``execute`` / ``get_operation`` are WorkerApi.Init, WorkerApi.Execute and
    LongRunningService.Get; `lzy_b200.grpc_api.server.serve(worker=...)` puts them on the wire."""

    def __init__(self, cm: H.ChannelManager, gpu: int = 0, api_url: str = "lzyb://worker-0", *,
                 mount_factory: Optional[MountFactory] = None, env_factory: Optional[EnvFactory] = None,
                 command_runner: Optional[CommandRunner] = None, vm_id: str = "", local_socket: Optional[str] = None):
        self.cm, self.gpu, self.api_url, self.vm_id = cm, gpu, api_url, vm_id or f"vm-{uuid.uuid4().hex[:8]}"
        # a unix socket on which this worker answers LzybPeerApi.Describe for consumers of the same box (csrc/host/
        # local_rpc.cpp); `api_url` must then carry "uds=<that path>" in its query
        self.local_socket = local_socket
        self._lock = threading.RLock()
        self._done = threading.Condition(self._lock)  # notified whenever an operation completes
        self._owner: Optional[Tuple[str, str]] = None
        self.slots: Optional[H.Slots] = None  # exists only after Init (WorkerApiImpl.java:255-264)
        self._active = threading.Event()      # hasActiveExecution
        self._ops: Dict[str, Operation] = {}
        self._by_idempotency_key: Dict[str, str] = {}
        self._contexts: Dict[str, List[H.ExecutionContext]] = {}
        self._threads: List[threading.Thread] = []
        self._custom_mount = mount_factory is not None
        self.mount_factory: MountFactory = mount_factory or self._frame_mount
        self.env_factory: EnvFactory = env_factory or (lambda spec: None)
        self.command_runner: CommandRunner = command_runner or run_startup_in_process
        self._store = None
        self._stream = None

    # ---- the HBM side: only touched when no test mount is installed ------------------------------------
    def _frame_mount(self, ctx: H.ExecutionContext):
        from ..api.v1.b200.store import FrameStore
        from ..api.v1.startup import FrameMount
        from ..devmem import current_stream_ptr
        if self._store is None:
            self._store = FrameStore(self.gpu)
        return FrameMount(ctx, self._store, self.gpu, current_stream_ptr(self.gpu))

    @contextlib.contextmanager
    def device_scope(self):
        """the op body and the packs of its results run on this worker's GPU and compute stream"""
        if self._custom_mount:
            yield
            return
        import torch
        if self._stream is None:
            self._stream = torch.cuda.Stream(self.gpu)
        with torch.cuda.device(self.gpu), torch.cuda.stream(self._stream):
            yield

    # ---- WorkerApi.Init ---------------------------------------------------------------------------------
    def init(self, user_id: str, workflow_name: str, worker_subject_name: str = "", worker_private_key: str = "") -> None:
        requester = (user_id, workflow_name)
        with self._lock:
            if self._owner is None:
                self._owner = requester
                try:
                    self.slots = H.Slots(self.cm, self.gpu, self.api_url)
                    if self.local_socket:
                        self.slots.serve_local(self.local_socket)
                except Exception as e:
                    self._owner = None
                    raise WorkerStatusError("INTERNAL", f"Cannot start slots server: {e}")
                return
            if self._owner != requester:
                _LOG.error("Attempt to execute op from another owner. Current is %s, got %s", self._owner, requester)
                raise WorkerStatusError("PERMISSION_DENIED")
        # the same owner reuses the worker for another execution: wait for the running task
        while self._active.is_set():
            _LOG.warning("Worker still has active execution, wait...")
            time.sleep(0.05)

    # ---- WorkerApi.Execute ------------------------------------------------------------------------------
    def execute(self, task: TaskSpec, idempotency_key: Optional[str] = None) -> Operation:
        with self._lock:
            requester = (task.user_id, task.workflow_name)
            if self._owner is None or self._owner != requester:
                _LOG.error("Attempt to execute op from another owner. Current is %s, got %s", self._owner, requester)
                raise WorkerStatusError("PERMISSION_DENIED")
            if self.slots is None:
                raise WorkerStatusError("INTERNAL")
            if idempotency_key and idempotency_key in self._by_idempotency_key:
                return self._ops[self._by_idempotency_key[idempotency_key]].snapshot()
            now = time.time()
            op = Operation(id=uuid.uuid4().hex, created_by=task.task_id, created_at=now, modified_at=now,
                           description=f"Worker, executionId: {task.execution_id}, taskId: {task.task_id}")
            self._ops[op.id] = op
            if idempotency_key:
                self._by_idempotency_key[idempotency_key] = op.id
            self._active.set()
            t = threading.Thread(target=self._run, args=(op, task), name=f"lzyb-task-{task.task_id}", daemon=True)
            self._threads.append(t)
            t.start()
            return op.snapshot()

    def _run(self, op: Operation, task: TaskSpec) -> None:
        try:
            result = self.execute_op(task)
            with self._lock:
                op.response, op.done, op.modified_at = result, True, time.time()
                self._active.clear()
                self._done.notify_all()
        except BaseException as e:  # LocalOperationService turns an escaped error into the operation's error
            with self._lock:
                op.error, op.done, op.modified_at = (13, f"{type(e).__name__}: {e}"), True, time.time()
                self._active.clear()
                self._done.notify_all()
        finally:
            self._active.clear()

    def execute_op(self, task: TaskSpec) -> ExecuteResult:
        """WorkerApiImpl.executeOp (:143-227)."""
        for slot in task.slots:
            if slot.name not in task.slot_assignments:
                _LOG.error("Empty binding for slot %s", slot.name)
                return ExecuteResult(ReturnCodes.INTERNAL_ERROR, "Internal error")
        _LOG.info("Configure worker...")
        try:
            self.env_factory(task.env)
        except InstallationError as e:
            _LOG.error("Unable to install environment: %s", e)
            return ExecuteResult(ReturnCodes.ENVIRONMENT_INSTALLATION_ERROR, str(e))
        except Exception:
            _LOG.exception("Error while preparing env")
            return ExecuteResult(ReturnCodes.INTERNAL_ERROR, "Internal error")
        _LOG.info("Executing task...")
        assert self.slots is not None
        mark = _phase_marker(task.task_id)
        try:
            ctx = self.slots.context(task.execution_id, task.task_id)
            with self._lock:
                self._contexts.setdefault(task.execution_id, []).append(ctx)
            for slot in task.slots:
                ctx.add_slot(slot.name, task.slot_assignments[slot.name], slot.direction)
            mark("slots added")
            _LOG.info("Waiting for slots...")
            ctx.before_execution(-1)
            mark("before_execution")
            with applied_env(_env_vars(task.env)):  # Operation.env.env: the variables the op asked for
                rc = self.command_runner(task.command, ctx, self)
            mark("command")
            if rc == 0:
                message = "Success"
                ctx.after_execution(-1)
                mark("after_execution")
            else:
                message = "Error while executing command on worker. See your stdout/stderr to see more info."
            return ExecuteResult(rc, message)
        except Exception:
            _LOG.exception("Error while executing task, stopping worker")
            return ExecuteResult(ReturnCodes.INTERNAL_ERROR, "Internal error")

    # ---- LongRunningService.Get -------------------------------------------------------------------------
    def get_operation(self, operation_id: str) -> Operation:
        with self._lock:
            op = self._ops.get(operation_id)
            if op is None:
                raise WorkerStatusError("NOT_FOUND", f"Operation {operation_id} not found")
            return op.snapshot()

    def wait_operation(self, operation_id: str, timeout: float = 60.0) -> Operation:
        """the operation as soon as it is done, or as it stands after `timeout` seconds — woken by the completing task, no
        polling (what a ``Get`` that carries a ``lzyb-wait-ms`` header is answered with, grpc_api/server.py)"""
        deadline = time.monotonic() + timeout
        with self._done:
            while True:
                op = self.get_operation(operation_id)
                left = deadline - time.monotonic()
                if op.done or left <= 0:
                    return op
                self._done.wait(left)

    # ---- lifecycle --------------------------------------------------------------------------------------
    def release_execution(self, execution_id: str) -> None:
        """drop the slots of a finished execution (the reference does this when the channel manager destroys
        the execution's channels and the worker is returned to the allocator)"""
        with self._lock:
            ctxs = self._contexts.pop(execution_id, [])
        for ctx in ctxs:
            ctx.close()

    def close(self) -> None:
        for t in list(self._threads):
            t.join(timeout=30)
        with self._lock:
            ctxs = [c for v in self._contexts.values() for c in v]
            self._contexts.clear()
        for ctx in ctxs:
            ctx.close()
        if self.slots is not None:
            self.slots.close()
            self.slots = None
        self._owner = None
