"""LzyWorkflow (pylzy/lzy/core/workflow.py:39-298): the call queue, the barrier, result-URI rules.

Only one workflow may be active per process (:205-215).  Calls are lazy: they run when a proxy is
touched, on ``barrier()`` or on exit; ``eager=True`` runs each call immediately (:118-124,154-171).
Result URIs (``cache=False``: ``<storage>/lzy_runs/<wf>/ops/<fn>/<call.id>/return_<i>``; ``cache=True``:
``<storage>/[<user>/]lzy_runs/<wf>/ops/<fn>/<version>/<input hashes>/return_<i>``) follow :247-298."""
from __future__ import annotations

import logging
import os
import re
from dataclasses import dataclass, replace
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
NAME_VALID_SYMBOLS = "letters, digits, '-', '_'"
_NAME_RE = re.compile(r"^[A-Za-z0-9_\-]+$")
_LOG = logging.getLogger("lzy_b200.workflow")


@dataclass
class LzyWorkflow(WithEnvironmentMixin):
    instance: ClassVar[Optional["LzyWorkflow"]] = None
    name: str
    owner: "Lzy"
    env: LzyEnvironment
    eager: bool = False
    interactive: bool = True

    def with_fields(self, **kwargs: Any) -> "LzyWorkflow":
        """a copy with some declaration changed — only before the workflow is used: once it has calls, a snapshot
        or an execution it carries state a copy could not share (pylzy/lzy/core/workflow.py:77-96)"""
        if self.__filled_entry_ids or self.__call_queue or self.__started or self.__snapshot or self.__execution_id:
            raise RuntimeError("It's forbidden to change LzyWorkflow fields after start working with it")
        return replace(self, **kwargs)

    @classmethod
    def get_active(cls) -> Optional["LzyWorkflow"]:
        if op_scope.inside_op():  # an op body sees only a workflow it opened itself, as in the reference's op subprocess
            return op_scope.get_workflow()
        return cls.instance

    @classmethod
    def set_active(cls, instance: Optional["LzyWorkflow"]) -> None:
        for klass in cls.__mro__:
            if isinstance(klass, type) and issubclass(klass, LzyWorkflow):
                klass.instance = instance

    def __post_init__(self) -> None:
        if not _NAME_RE.match(self.name or ""):
            raise ValueError(f"Invalid workflow name. Name can contain only {NAME_VALID_SYMBOLS}")
        self.__user = os.getenv(USER_ENV)
        self.__filled_entry_ids: Set[str] = set()
        self.__call_queue: List["LzyCall"] = []
        self.__started = False
        self.__nested = False
        self.__snapshot: Optional[Snapshot] = None
        self.__execution_id: Optional[str] = None

    @property
    def log(self) -> logging.Logger:
        return _LOG

    @property
    def snapshot(self) -> Snapshot:
        if self.__snapshot is None:
            raise ValueError("Workflow is not yet started")
        return self.__snapshot

    @property
    def execution_id(self) -> str:
        if self.__execution_id is None:
            raise ValueError("Workflow is not yet started")
        return self.__execution_id

    @property
    def filled_entry_ids(self) -> Set[str]:
        return self.__filled_entry_ids

    @property
    def call_queue(self) -> List["LzyCall"]:
        return self.__call_queue

    def register_call(self, call: "LzyCall") -> Any:
        self.__call_queue.append(call)
        if self.eager:
            self.barrier()

    def barrier(self) -> None:
        LzyEventLoop.run_async(self._barrier())

    def __enter__(self) -> "LzyWorkflow":
        try:
            parts = [self.owner.storage_uri]
            if self.__user is not None:
                parts.append(self.__user)
            parts.extend(["lzy_runs", self.name, "inputs/"])
            storage_uri = "/".join(parts)
            # the op thread that opens a workflow owns it: decided here, in the caller's thread (the coroutine
            # below runs on the workflow thread)
            self.__nested = op_scope.inside_op()
            if self.get_active() is not None:
                raise RuntimeError("Simultaneous workflows are not supported")
            self.__execution_id = LzyEventLoop.run_async(self.__start())
            self.__activate(self)
            self.__snapshot = self.owner.make_snapshot(self, storage_uri)
            return self
        except Exception:
            try:
                self.__abort()
            finally:
                raise

    def __exit__(self, exc_type, exc_val, exc_tb) -> None:
        try:
            if not self.__started:
                self.log.warning("Exiting workflow that has not been started")
                return
            if exc_val is None:
                try:
                    self.barrier()
                except Exception as e:  # pylint: disable=broad-exception-caught
                    exc_val = e
        finally:
            if exc_val is None:
                self.__destroy()
            else:
                try:
                    self.__abort()
                finally:
                    raise exc_val

    def __activate(self, workflow: Optional["LzyWorkflow"]) -> None:
        if self.__nested:
            op_scope.set_workflow(workflow)
        else:
            self.set_active(workflow)

    def __abort(self):
        if not self.__started:
            # nothing of this workflow is running: the runtime is left alone (pylzy/tests/api/v1/test_remote_runtime.py
            # :65-76, "should NOT be aborted if not started") — in particular a refused second workflow does not
            # take down the one that is active
            return
        try:
            LzyEventLoop.run_async(self.owner.runtime.abort())
        finally:
            self.__activate(None)
            self.__started = False

    def __destroy(self):
        try:
            LzyEventLoop.run_async(self.owner.runtime.finish())
        finally:
            self.__activate(None)
            self.__started = False

    async def __start(self) -> str:
        if self.__started:
            raise RuntimeError("Workflow already started")
        workflow_id = await self.owner.runtime.start(self)
        self.__started = True
        return workflow_id

    async def _barrier(self) -> None:
        if not self.__call_queue:
            return
        for call in self.__call_queue:
            if call.cache:
                self.__gen_results_uri_in_cache(call)
            else:
                self.__gen_results_uri_skip_cache(call)
        queue, self.__call_queue = self.__call_queue, []
        try:
            await self.owner.runtime.exec(queue, lambda x: self.log.info(f"Graph status: {x.name}"))
        except BaseException:
            # the reference leaves the queue intact when exec raises (the assignment below it is never
            # reached, workflow.py:230-234) and the workflow is aborted by __exit__; same net effect
            self.__call_queue = []
            raise

    def __gen_results_uri_in_cache(self, call: "LzyCall"):
        args_hashes = [self.snapshot.get(eid).data_hash for eid in call.arg_entry_ids]
        kwargs_hashes = [f"{name}:{self.snapshot.get(call.kwarg_entry_ids[name]).data_hash}" for name in sorted(call.kwargs.keys())]
        inputs_hashes_concat = "_".join([*args_hashes, *kwargs_hashes])
        parts = [self.owner.storage_uri, self.__user or "", "lzy_runs", self.name, "ops", call.callable_name, call.version,
                 inputs_hashes_concat]
        uri_prefix = "/".join(filter(bool, parts))
        for i, eid in enumerate(call.entry_ids):
            if eid not in self.__filled_entry_ids:
                entry = self.snapshot.get(eid)
                entry.storage_uri = uri_prefix + f"/return_{i}"
                entry.data_hash = md5_of_str(entry.storage_uri)
                self.__filled_entry_ids.add(eid)
        self.__fill_exception_entry(call)

    def __gen_results_uri_skip_cache(self, call: "LzyCall"):
        for i, eid in enumerate(call.entry_ids):
            if eid not in self.__filled_entry_ids:
                entry = self.snapshot.get(eid)
                entry.storage_uri = f"{self.owner.storage_uri}/lzy_runs/{self.name}/ops/{call.callable_name}/{call.id}/return_{i}"
                entry.data_hash = md5_of_str(entry.storage_uri)
                self.__filled_entry_ids.add(eid)
        self.__fill_exception_entry(call)

    def __fill_exception_entry(self, call: "LzyCall"):
        eid = call.exception_id
        if eid not in self.__filled_entry_ids:
            entry = self.snapshot.get(eid)
            entry.storage_uri = f"{self.owner.storage_uri}/lzy_runs/{self.name}/ops/{call.callable_name}/{call.id}/exception"
            entry.data_hash = md5_of_str(entry.storage_uri)
            self.__filled_entry_ids.add(eid)
