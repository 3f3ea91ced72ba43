"""lzy_b200 — a B200-native slot/channel data plane behind the pylzy ``@op`` / ``Lzy.workflow`` API.

    from lzy_b200 import Lzy, op            # same names as `from lzy.api.v1 import Lzy, op`

The data plane itself is the C ABI in ``include/lzy_b200.h`` (``lzy_b200._native``); the host-side
control logic (channel manager, slot state machines, graph dispatch) is ``include/lzy_b200_host.h``
(``lzy_b200._host``).  There is no CPU data path.
"""
import importlib
import importlib.util
import sys

__version__ = "0.1.0"

_LAZY = {"Lzy", "op", "LzyWorkflow", "B200Runtime", "B200RuntimeConfig", "Runtime", "materialize", "materialized",
         "is_lzy_proxy", "LzyEnvironment", "Provisioning"}


def __getattr__(name):
    if name in _LAZY:
        from .api import v1
        return getattr(v1, name)
    raise AttributeError(name)


class _LzyAlias:
    """meta-path finder + loader: ``import lzy.a.b`` yields the very module object of ``lzy_b200.a.b`` (not a
    second copy of it), so classes pickled by reference under either name are the same classes"""

    def find_spec(self, fullname, path=None, target=None):
        if fullname != "lzy" and not fullname.startswith("lzy."):
            return None
        real = __name__ + fullname[3:]
        try:
            importlib.import_module(real)
        except ModuleNotFoundError as e:
            if e.name and real.startswith(e.name):
                return None
            raise
        return importlib.util.spec_from_loader(fullname, self, is_package=hasattr(sys.modules[real], "__path__"))

    def create_module(self, spec):
        return sys.modules[__name__ + spec.name[3:]]

    def exec_module(self, module):
        pass


def install_as_lzy() -> None:
    """Alias this package as ``lzy`` so that unmodified user code (`from lzy.api.v1 import Lzy, op`) runs on the
    B200 data plane, and objects a stock pylzy client pickled by reference (``lzy.api.v1.startup.ProcessingRequest``)
    resolve to the classes of this package."""
    if not any(isinstance(f, _LzyAlias) for f in sys.meta_path):
        sys.meta_path.insert(0, _LzyAlias())
    sys.modules.setdefault("lzy", sys.modules[__name__])
