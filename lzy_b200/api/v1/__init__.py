"""``from lzy_b200.api.v1 import Lzy, op`` — the user-facing names of ``lzy.api.v1``
(pylzy/lzy/api/v1/__init__.py)."""
from ...core.call import LazyCallWrapper, LzyCall
from ...core.lzy import Lzy
from ...core.op import op
from ...core.workflow import LzyWorkflow
from ...env import LzyEnvironment, Provisioning, env_vars, provisioning
from .b200 import B200Runtime, B200RuntimeConfig
from .exceptions import LzyExecutionException
from .runtime import ProgressStep, Runtime
from .snapshot import Snapshot, SnapshotEntry
from .utils.proxy_adapter import is_lzy_proxy, materialize, materialized

__all__ = ["Lzy", "op", "LzyWorkflow", "LzyCall", "LazyCallWrapper", "Runtime", "ProgressStep", "B200Runtime",
           "B200RuntimeConfig", "Snapshot", "SnapshotEntry", "LzyEnvironment", "Provisioning",
           "LzyExecutionException", "is_lzy_proxy", "materialize", "materialized", "env_vars", "provisioning"]
