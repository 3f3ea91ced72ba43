"""lzy_b200 — a B200-native slot/channel data plane behind the pylzy ``@op`` / ``Lzy.workflow`` API.

    from lzy_b200 import Lzy, op            # same names as `from lzy.api.v1 import Lzy, op`

The data plane itself is the C ABI in ``include/lzy_b200.h`` (``lzy_b200._native``); the host-side
control logic (channel manager, slot state machines, graph dispatch) is ``include/lzy_b200_host.h``
(``lzy_b200._host``).  There is no CPU data path.
"""
__version__ = "0.1.0"

_LAZY = {"Lzy", "op", "LzyWorkflow", "B200Runtime", "B200RuntimeConfig", "Runtime", "materialize", "materialized",
         "is_lzy_proxy", "LzyEnvironment", "Provisioning"}


def __getattr__(name):
    if name in _LAZY:
        from .api import v1
        return getattr(v1, name)
    raise AttributeError(name)


def install_as_lzy() -> None:
    """Alias this package as ``lzy`` so that unmodified user code (`from lzy.api.v1 import Lzy, op`)
    runs on the B200 data plane."""
    import sys
    from . import api, types
    from .api import v1
    sys.modules.setdefault("lzy", sys.modules[__name__])
    sys.modules.setdefault("lzy.api", api)
    sys.modules.setdefault("lzy.api.v1", v1)
    sys.modules.setdefault("lzy.types", types)
