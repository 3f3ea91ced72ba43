"""The ``Runtime`` seam: what a workflow needs from whatever runs its ops.  The five coroutines and
``ProgressStep`` are the interface of pylzy/lzy/api/v1/runtime.py:21-44 — this is the boundary a runtime written
for stock pylzy plugs into (``Lzy(runtime=...)``, pylzy/lzy/core/lzy.py:48), so names and signatures are kept
as they are there.  All of them run on the workflow thread (``LzyEventLoop``).

Call order, as ``LzyWorkflow`` drives it (pylzy/lzy/core/workflow.py:173-236):

    start(workflow)                    once, on ``__enter__``; its answer is the execution id
    exec(calls, progress) …            once per barrier, with the queued calls in submission order
    finish()  |  abort()               once, on ``__exit__``: finish after a clean block, abort after any exception

A runtime that keeps several executions apart (an op body may open a workflow of its own on the same runtime) can
additionally offer ``finish_workflow(workflow)`` / ``abort_workflow(workflow)``; ``LzyWorkflow`` calls those instead of
``finish()`` / ``abort()`` when they exist.  Stock runtimes need not know about them.

``snapshot()`` is this package's other addition: the workflow asks its runtime for the ``Snapshot`` to keep entries
in, where the reference hard-wires ``DefaultSnapshot`` over a storage client (pylzy/lzy/core/workflow.py:133-146;
the alternative seam named in pylzy/lzy/api/v1/snapshot.py:55-74).  The HBM runtime answers with a snapshot whose
entries are frames in device memory."""
from __future__ import annotations

import enum
from abc import ABC, abstractmethod
from typing import TYPE_CHECKING, Any, Callable, List, Optional

if TYPE_CHECKING:
    from ...core.call import LzyCall
    from ...core.workflow import LzyWorkflow
    from .snapshot import Snapshot


class ProgressStep(enum.Enum):
    """what ``exec`` reports through its ``progress`` callback while a graph runs"""
    WAITING = "WAITING"
    EXECUTING = "EXECUTING"
    COMPLETED = "COMPLETED"
    FAILED = "FAILED"


class Runtime(ABC):
    @abstractmethod
    async def start(self, workflow: "LzyWorkflow") -> str:
        """begin an execution of `workflow`; returns its id"""

    @abstractmethod
    async def exec(self, calls: List["LzyCall"], progress: Callable[[ProgressStep], None]) -> None:
        """run `calls` (a valid topological order of one graph).  Afterwards every ``call.entry_ids[i]`` and every
        ``call.exception_id`` is readable through ``workflow.snapshot.get_data``.  When an op raises, its own
        exception — type, arguments, traceback — is raised from here."""

    @abstractmethod
    async def finish(self) -> None:
        """the execution ended normally; a runtime that never started has nothing to do"""

    @abstractmethod
    async def abort(self) -> None:
        """the execution is given up (an exception left the ``with`` block); same no-op rule as ``finish``"""

    @abstractmethod
    async def storage(self) -> Optional[Any]:
        """the storage the service side assigned to this user, if the runtime has one to offer"""

    def snapshot(self, workflow: "LzyWorkflow", storage_uri_prefix: str) -> Optional["Snapshot"]:
        """the ``Snapshot`` for `workflow`'s entries (local data goes under `storage_uri_prefix`); None = the runtime
        has none to offer, which ``Lzy.make_snapshot`` reports as an error in this build"""
        return None
