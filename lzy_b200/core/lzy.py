"""The ``Lzy`` entry object (pylzy/lzy/core/lzy.py:46-176): runtime + serializer registry + storage
identity; ``workflow(name, eager=..., interactive=...)`` opens a workflow."""
from __future__ import annotations

import sys
from dataclasses import dataclass, field, replace
from typing import Any, ClassVar, Optional, Type

from ..api.v1.runtime import Runtime
from ..api.v1.snapshot import Snapshot
from ..env import LzyEnvironment, WithEnvironmentMixin
from ..serialization.registry import LzySerializerRegistry
from .workflow import LzyWorkflow


def _default_runtime() -> Runtime:
    from ..api.v1.b200.runtime import B200Runtime
    return B200Runtime()


@dataclass(frozen=True)
class Lzy(WithEnvironmentMixin):
    env: LzyEnvironment = field(default_factory=LzyEnvironment)
    runtime: Runtime = field(default_factory=_default_runtime)
    serializer_registry: LzySerializerRegistry = field(default_factory=LzySerializerRegistry)
    storage_uri: str = "hbm://lzy_b200"
    storage_name: str = "hbm"

    _workflow_class: ClassVar[Type[LzyWorkflow]] = LzyWorkflow

    def with_fields(self, **kwargs: Any) -> "Lzy":
        return replace(self, **kwargs)

    def workflow(self, name: str, *, eager: bool = False, interactive: bool = True,
                 env: Optional[LzyEnvironment] = None) -> LzyWorkflow:
        frame = sys._getframe(1)  # the caller's frame; inspect.stack() would also read source lines of every frame
        namespace = {**frame.f_globals, **frame.f_locals}
        env = env or LzyEnvironment()
        env = env.with_fields(namespace={**namespace, **env.get_namespace()})
        return self._workflow_class(name=name, owner=self, env=env, eager=eager, interactive=interactive)

    def make_snapshot(self, workflow: LzyWorkflow, storage_uri_prefix: str) -> Snapshot:
        snap = self.runtime.snapshot(workflow, storage_uri_prefix)
        if snap is None:
            raise RuntimeError(
                f"{type(self.runtime).__name__}.snapshot() returned None: this build keeps workflow data in HBM "
                f"frames and has no file/S3 storage clients; supply a Snapshot from the runtime")
        return snap
