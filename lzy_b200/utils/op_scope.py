"""Thread-local marker set while an op body runs on a worker thread.  Inside it there is no active
workflow — exactly as inside the reference's op subprocess — so nested ``@op`` calls are plain calls."""
import threading

_scope = threading.local()


def enter() -> None:
    _scope.depth = getattr(_scope, "depth", 0) + 1


def leave() -> None:
    _scope.depth = getattr(_scope, "depth", 1) - 1


def inside_op() -> bool:
    return bool(getattr(_scope, "depth", 0))
