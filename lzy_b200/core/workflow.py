"""``LzyWorkflow`` — the ``with lzy.workflow(...)`` object (behaviour of pylzy/lzy/core/workflow.py:39-298).

* One workflow is active per process (:205-215); an op body — a thread here, a process of its own in the
  reference — sees none of its caller's and may open its own.
* ``@op`` calls are queued and run by ``barrier()``: when a proxy is touched, on request, or on leaving the
  ``with`` block; an eager workflow runs each call as it is registered (:118-124, :154-171).
* Just before a queue is handed to the runtime every result gets its storage URI (:247-298):
  ``cache=False``  ``<storage>/lzy_runs/<workflow>/ops/<function>/<call id>/return_<i>``
  ``cache=True``   ``<storage>/[<user>/]lzy_runs/<workflow>/ops/<function>/<version>/<argument hashes>/return_<i>``
  (hashes joined by ``_``, keyword arguments as ``name:hash`` sorted by name — equal calls meet at equal URIs, which
  is all that caching is); the exception of a call always lives under its call id.
* Leaving the block cleanly finishes the execution; any exception (the block's own, or one from the last barrier)
  aborts it and is re-raised."""
from __future__ import annotations

import logging
import os
import re
from typing import TYPE_CHECKING, Any, ClassVar, List, Optional, Set

from ..api.v1.snapshot import Snapshot
from ..api.v1.utils.hashing import md5_of_str
from ..env import LzyEnvironment, WithEnvironmentMixin
from ..utils import op_scope
from ..utils.event_loop import LzyEventLoop

if TYPE_CHECKING:
    from .call import LzyCall
    from .lzy import Lzy

USER_ENV = "LZY_USER"
NAME_VALID_SYMBOLS = "[alphanumeric characters,_,-]"  # the wording of pylzy/lzy/api/v1/utils/validation.py:3
# what `re.match("^[a-zA-Z0-9_-]*$", name)` accepts there (:6-7) — the empty name and one trailing newline included
_VALID_NAME = re.compile(r"[A-Za-z0-9_-]*\n?")
_LOG = logging.getLogger("lzy_b200.workflow")


class LzyWorkflow(WithEnvironmentMixin):
    instance: ClassVar[Optional["LzyWorkflow"]] = None

    def __init__(self, name: str, owner: "Lzy", env: LzyEnvironment, eager: bool = False, interactive: bool = True):
        if not isinstance(name, str) or not _VALID_NAME.fullmatch(name):
            raise ValueError(f"Invalid workflow name. Name can contain only {NAME_VALID_SYMBOLS}")
        self.name, self.owner, self.env, self.eager, self.interactive = name, owner, env, eager, interactive
        self._user = os.getenv(USER_ENV)
        self._queue: List["LzyCall"] = []
        self._uri_assigned: Set[str] = set()   # entries whose storage URI has been decided
        self._snapshot: Optional[Snapshot] = None
        self._execution: Optional[str] = None
        self._running = False
        self._start_attempted = False
        self._opened_inside_op = False

    # ---- declarations ---------------------------------------------------------------------------------------------
    def with_fields(self, **changes: Any) -> "LzyWorkflow":
        """a copy with some declaration changed.  Only an untouched workflow can be copied (:77-96): once it has
        calls, a snapshot or an execution it carries state a copy could not share."""
        if self._running or self._queue or self._uri_assigned or self._snapshot is not None or self._execution is not None:
            raise RuntimeError("It's forbidden to change LzyWorkflow fields after start working with it")
        declared = dict(name=self.name, owner=self.owner, env=self.env, eager=self.eager, interactive=self.interactive)
        declared.update(changes)
        return type(self)(**declared)

    def __repr__(self) -> str:
        return f"{type(self).__name__}(name={self.name!r}, eager={self.eager}, interactive={self.interactive})"

    # ---- which workflow is active ---------------------------------------------------------------------------------
    @classmethod
    def get_active(cls) -> Optional["LzyWorkflow"]:
        if op_scope.inside_op():
            return op_scope.get_workflow()  # only a workflow this op body opened itself
        return cls.instance

    @classmethod
    def set_active(cls, instance: Optional["LzyWorkflow"]) -> None:
        # subclasses (Lzy._workflow_class) and the base class answer get_active() alike
        for klass in cls.__mro__:
            if isinstance(klass, type) and issubclass(klass, LzyWorkflow):
                klass.instance = instance

    def _mark_active(self, on: bool) -> None:
        if self._opened_inside_op:
            op_scope.set_workflow(self if on else None)
        else:
            self.set_active(self if on else None)

    # ---- state the rest of the API reads ------------------------------------------------------------------------
    @property
    def log(self) -> logging.Logger:
        return _LOG

    @property
    def snapshot(self) -> Snapshot:
        if self._snapshot is None:
            raise ValueError("Workflow is not yet started")
        return self._snapshot

    @property
    def execution_id(self) -> str:
        if self._execution is None:
            raise ValueError("Workflow is not yet started")
        return self._execution

    @property
    def filled_entry_ids(self) -> Set[str]:
        return self._uri_assigned

    @property
    def call_queue(self) -> List["LzyCall"]:
        return self._queue

    # ---- the call queue -------------------------------------------------------------------------------------------
    def register_call(self, call: "LzyCall") -> None:
        self._queue.append(call)
        if self.eager:
            self.barrier()

    def _inline(self) -> bool:
        """may the runtime's and the snapshot's coroutines be driven on the caller's thread? (LzyEventLoop.run_async)"""
        return bool(getattr(self.owner.runtime, "inline_coroutines", False)) and \
            (self._snapshot is None or bool(getattr(self._snapshot, "inline_coroutines", False)))

    def barrier(self) -> None:
        LzyEventLoop.run_async(self._barrier(), inline=self._inline())

    async def _barrier(self) -> None:
        if not self._queue:
            return
        for call in self._queue:
            self._assign_uris(call)
        submitted, self._queue = self._queue, []  # whatever happens to this graph, it is not submitted twice
        await self.owner.runtime.exec(submitted, lambda step: self.log.info(f"Graph status: {step.name}"))

    # ---- entering and leaving -----------------------------------------------------------------------------------
    def __enter__(self) -> "LzyWorkflow":
        # whose workflow this is gets decided here, in the caller's thread: the coroutines run on the workflow thread
        self._opened_inside_op = op_scope.inside_op()
        try:
            if self._running:
                raise RuntimeError("Workflow already started")
            if self.get_active() is not None:
                raise RuntimeError("Simultaneous workflows are not supported")
            self._start_attempted = True
            self._execution = LzyEventLoop.run_async(self.owner.runtime.start(self), inline=self._inline())
            self._running = True
            self._mark_active(True)
            where = [self.owner.storage_uri] + ([] if self._user is None else [self._user])
            self._snapshot = self.owner.make_snapshot(self, "/".join(where + ["lzy_runs", self.name, "inputs/"]))
        except Exception:
            self._leave(aborting=True)
            raise
        return self

    def __exit__(self, exc_type, exc_val, exc_tb) -> None:
        if not self._running:
            self.log.warning("Exiting workflow that has not been started")
            return
        failure = exc_val
        if failure is None:
            try:
                self.barrier()
            except Exception as e:  # pylint: disable=broad-exception-caught
                failure = e
        try:
            self._leave(aborting=failure is not None)
        finally:
            if failure is not None:
                raise failure

    def _leave(self, aborting: bool) -> None:
        """finish() after a clean block, abort() otherwise — also when ``start`` itself raised, as the reference does
        (its runtime decides what an abort without a start means, pylzy/tests/api/v1/test_remote_runtime.py:65-76).
        A workflow that was *refused* before it asked the runtime for anything — another one is active — leaves the
        runtime and the active workflow alone; the reference aborts the running execution in that case."""
        if not self._start_attempted:
            return
        runtime, started = self.owner.runtime, self._running
        # a runtime that keeps several executions apart (nested workflows) is told whose end this is
        by_workflow = getattr(runtime, "abort_workflow" if aborting else "finish_workflow", None)
        try:
            if by_workflow is not None:
                LzyEventLoop.run_async(by_workflow(self), inline=self._inline())
            else:
                LzyEventLoop.run_async(runtime.abort() if aborting else runtime.finish(), inline=self._inline())
        finally:
            if started:
                self._mark_active(False)
            self._running = self._start_attempted = False

    # ---- storage URIs of results ---------------------------------------------------------------------------------
    def _assign_uris(self, call: "LzyCall") -> None:
        storage, ops_dir = self.owner.storage_uri, f"lzy_runs/{self.name}/ops/{call.callable_name}"
        per_call = f"{storage}/{ops_dir}/{call.id}"
        if call.cache:
            hash_of = lambda eid: self.snapshot.get(eid).data_hash  # noqa: E731
            inputs = [hash_of(eid) for eid in call.arg_entry_ids]
            inputs += [f"{name}:{hash_of(call.kwarg_entry_ids[name])}" for name in sorted(call.kwargs)]
            results_at = "/".join(filter(None, [storage, self._user, ops_dir, call.version, "_".join(inputs)]))
        else:
            results_at = per_call
        for i, eid in enumerate(call.entry_ids):
            self._place(eid, f"{results_at}/return_{i}")
        self._place(call.exception_id, f"{per_call}/exception")

    def _place(self, entry_id: str, uri: str) -> None:
        if entry_id in self._uri_assigned:
            return
        entry = self.snapshot.get(entry_id)
        entry.storage_uri = uri
        entry.data_hash = md5_of_str(uri)
        self._uri_assigned.add(entry_id)
