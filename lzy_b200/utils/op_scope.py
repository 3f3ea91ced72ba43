"""Thread-local marker set while an op body runs on a worker thread.  Inside it the caller's workflow is not
visible — exactly as inside the reference's op subprocess — so nested ``@op`` calls are plain calls, and an op
may open a workflow of its own (pylzy/tests/scenarios/nested_workflows), which then belongs to that op's thread."""
import threading

_scope = threading.local()


def enter() -> None:
    _scope.depth = getattr(_scope, "depth", 0) + 1


def leave() -> None:
    _scope.depth = getattr(_scope, "depth", 1) - 1


def inside_op() -> bool:
    return bool(getattr(_scope, "depth", 0))


def get_workflow():
    """the workflow opened *inside* the op running on this thread, if any"""
    return getattr(_scope, "workflow", None)


def set_workflow(workflow) -> None:
    _scope.workflow = workflow
