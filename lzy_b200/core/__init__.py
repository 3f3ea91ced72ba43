"""``lzy.core`` — workflows, calls and the ``@op`` decorator.  This package and ``lzy_b200.api.v1`` import each other
(as their namesakes in pylzy do); entering through the API package first makes ``import lzy_b200.core.call`` work on
its own as well."""
from ..api import v1 as _api_v1  # noqa: F401
