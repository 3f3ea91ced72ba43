"""The reference's workflow scenarios (pylzy/tests/core/test_workflow.py — the assertions that must hold for any
`Runtime`) on a box without a GPU: the very test functions of tests/test_workflow_gpu.py, run against
`WorkerPoolRuntime` (tests/support/worker_pool_runtime.py), which composes this repository's graph dispatcher,
channel manager, C++ slots runtime, worker front and `startup` entry point over the in-memory slot backend.
Scenarios that are about HBM frames or CUDA tensors stay GPU-only."""
import uuid

import pytest

from lzy_b200.api.v1 import Lzy, op
from lzy_b200.api.v1.utils.proxy_adapter import materialize
from tests import test_workflow_gpu as scenarios
from tests.support.worker_pool_runtime import WorkerPoolRuntime

_RUNTIME = None


@pytest.fixture(scope="module", autouse=True)
def _teardown_runtime():
    yield
    global _RUNTIME
    if _RUNTIME is not None:
        _RUNTIME.close()
        _RUNTIME = None


@pytest.fixture
def lzy():
    global _RUNTIME
    if _RUNTIME is None:
        _RUNTIME = WorkerPoolRuntime(n_workers=2)
    return Lzy(runtime=_RUNTIME, storage_uri="mem://bucket/prefix")


@pytest.fixture
def workflow_name():
    return f"test_workflow_{uuid.uuid4().hex}"


# the same function objects, collected here without the module-level `gpu` mark
test_lists = scenarios.test_lists
test_invalid_list_type_returns = scenarios.test_invalid_list_type_returns
test_tuple_types = scenarios.test_tuple_types
test_optional_return = scenarios.test_optional_return
test_optional_and_union_args = scenarios.test_optional_and_union_args
test_globals_not_materialized = scenarios.test_globals_not_materialized
test_kwargs = scenarios.test_kwargs
test_return_accept_list = scenarios.test_return_accept_list
test_empty_and_unspecified_containers = scenarios.test_empty_and_unspecified_containers
test_lazy = scenarios.test_lazy
test_eager = scenarios.test_eager
test_materialize = scenarios.test_materialize
test_lazy_and_eager_args_loading = scenarios.test_lazy_and_eager_args_loading
test_local_startup_import = scenarios.test_local_startup_import
test_barrier_and_complex_graph = scenarios.test_barrier_and_complex_graph
test_exception_is_reraised_with_original_type_args_and_traceback = scenarios.test_exception_is_reraised_with_original_type_args_and_traceback
test_op_result_changed_by_ref_and_not_inside_op = scenarios.test_op_result_changed_by_ref_and_not_inside_op


def test_cached_op_runs_once_and_equal_local_data_is_one_blob(lzy, workflow_name):
    @op(cache=True, version="1")
    def expensive(x: int) -> int:
        return x * 2

    for _ in range(2):
        with lzy.workflow(workflow_name) as wf:
            assert 42 == expensive(21)
            uris = sorted(wf.snapshot.uris())
    assert lzy.runtime.executed.count("expensive") == 1 and lzy.runtime.skipped_cached >= 1
    assert len(uris) == len(set(uris))


def test_tasks_are_spread_over_the_workers_and_arrive_slot_to_slot(lzy, workflow_name):
    @op
    def inc(x: int) -> int:
        return x + 1

    for w in lzy.runtime.workers or []:
        if w.slots is not None:
            w.slots.transfer_log(clear=True)
    with lzy.workflow(workflow_name):
        y = inc(inc(inc(inc(0))))
        assert materialize(y) == 4
        log = [r for w in lzy.runtime.workers for r in w.slots.transfer_log()]
    assert len(log) == 4 and all(r["ok"] for r in log)
    # hop 1 comes from storage (the local argument), every later hop from the previous task's output slot
    assert sorted(r["peer_kind"] for r in log) == ["slot", "slot", "slot", "storage"]


def test_diamond_runs_in_dependency_order_and_a_failure_stops_the_graph(lzy, workflow_name):
    @op
    def src() -> str:
        return "s"

    @op
    def branch(t: str, k: int) -> str:
        return t + str(k)

    @op
    def bad(t: str) -> str:
        raise RuntimeError("branch failed")

    @op
    def sink(a: str, b: str) -> str:
        return a + "|" + b

    before = len(lzy.runtime.executed)
    with lzy.workflow(workflow_name):
        s = src()
        out = sink(branch(s, 1), branch(s, 2))
        assert materialize(out) == "s1|s2"
    order = lzy.runtime.executed[before:]
    assert order[0] == "src" and order[-1] == "sink" and sorted(order[1:3]) == ["branch", "branch"]

    before = len(lzy.runtime.executed)
    with pytest.raises(RuntimeError, match="branch failed"):
        with lzy.workflow(workflow_name + "_f"):
            s = src()
            out = sink(branch(s, 1), bad(s))
            materialize(out)
    assert "sink" not in lzy.runtime.executed[before:]  # its producers never both succeeded
