"""The ``Runtime`` seam (pylzy/lzy/api/v1/runtime.py:21-44): what a workflow needs from whatever runs
its ops.  ``Lzy(runtime=...)`` injects it (pylzy/lzy/core/lzy.py:48); ``LzyWorkflow`` calls ``start``
on entry, ``exec`` on every barrier and ``finish``/``abort`` on exit
(pylzy/lzy/core/workflow.py:173-236)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum
from typing import TYPE_CHECKING, Any, Callable, List, Optional

if TYPE_CHECKING:
    from ...core.call import LzyCall
    from ...core.workflow import LzyWorkflow
    from .snapshot import Snapshot


class ProgressStep(Enum):
    WAITING = "WAITING"
    EXECUTING = "EXECUTING"
    COMPLETED = "COMPLETED"
    FAILED = "FAILED"


class Runtime(ABC):
    @abstractmethod
    async def start(self, workflow: "LzyWorkflow") -> str:
        """returns the execution id"""

    @abstractmethod
    async def exec(self, calls: List["LzyCall"], progress: Callable[[ProgressStep], None]) -> None:
        """must leave every ``call.entry_ids[i]`` and ``call.exception_id`` readable through
        ``workflow.snapshot.get_data``; re-raises the op's own exception when an op fails"""

    @abstractmethod
    async def abort(self) -> None:
        ...

    @abstractmethod
    async def finish(self) -> None:
        ...

    @abstractmethod
    async def storage(self) -> Optional[Any]:
        ...

    # Not part of the reference ABC: lets a runtime supply the workflow's Snapshot (the alternative
    # seam named in pylzy/lzy/api/v1/snapshot.py:55-74).  The reference hard-wires DefaultSnapshot over
    # an AsyncStorageClient (pylzy/lzy/core/workflow.py:133-146); the B200 runtime keeps entries in HBM.
    def snapshot(self, workflow: "LzyWorkflow", storage_uri_prefix: str) -> Optional["Snapshot"]:
        return None
