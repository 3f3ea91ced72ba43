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


# ---- the reference's end-to-end scenarios (pylzy/tests/scenarios/*; PyApiTest.java:17-138 diffs their stdout) -------
def test_scenario_complex_graph_env_vars_and_none(lzy, workflow_name, capsys):
    """complex_graph/__init__.py:6-28, expected output `39 40 41 a=1,b=1 kek None 42`"""
    import os

    @op
    def repeatable_op(a: int) -> int:
        print(a)
        return a + 1

    @op
    def just_return_none(a: int, b: int) -> None:
        print(f"a={a},b={b}")
        print(os.environ["LOL"])
        return None

    assert "LOL" not in os.environ
    with lzy.workflow(workflow_name, interactive=False).with_env_vars({"LOL": "kek"}):
        a = 39
        for _ in range(3):
            a = repeatable_op(a)
        res = just_return_none(1, 1)
        print(res)
        print(a)
    words = capsys.readouterr().out.split()
    # the independent op may run beside the chain on the other worker: same lines, chain order kept
    assert sorted(words) == sorted(["None", "39", "40", "41", "a=1,b=1", "kek", "42"])
    assert [w for w in words if w in ("39", "40", "41", "42")] == ["39", "40", "41", "42"]
    assert words.index("a=1,b=1") + 1 == words.index("kek") and words[0] == "None"
    assert "LOL" not in os.environ  # restored after the op


def test_env_vars_combine_lzy_then_workflow_then_op(workflow_name):
    import os

    from tests.support.worker_pool_runtime import WorkerPoolRuntime

    @op
    def read() -> str:
        return ",".join(os.environ.get(k, "-") for k in ("A", "B", "C"))

    rt = WorkerPoolRuntime(n_workers=1)
    try:
        lzy = Lzy(runtime=rt, storage_uri="mem://bucket/env").with_env_vars(A="lzy", B="lzy", C="lzy")
        with lzy.workflow(workflow_name).with_env_vars(B="wf", C="wf"):
            assert read.with_env_vars({"C": "op"})() == "lzy,wf,op"
            assert read() == "lzy,wf,wf"
    finally:
        rt.close()


def test_scenario_exception_reaches_the_caller_and_is_not_cached(lzy, workflow_name, capsys):
    """exception_serialize, exec_fail and cached_exception: the op's own exception type comes back, and a failed
    cache=True op runs again next time (there is no result to reuse)"""
    @op(cache=True, version="1.0")
    def raises() -> None:
        print("exception was raised")
        raise ValueError("test")

    for _ in range(2):
        with pytest.raises(ValueError, match="test"):
            with lzy.workflow(workflow_name, interactive=False):
                raises()
    assert capsys.readouterr().out.count("exception was raised") == 2


def test_scenario_repeated_ops_and_executions_use_the_cache(lzy, workflow_name, capsys):
    """repeated_execs_use_cache + fully_cached_graph: `foo was called` once, bar (cache=False) every time, a fully
    cached second execution starts no task at all"""
    @op(cache=True, version="1.1")
    def foo_with_print(name: str, value: int) -> str:
        print("foo was called")
        return f"{name} is {value}"

    @op(cache=False)
    def bar_with_print(message: str) -> str:
        return f"message from bar: {message}"

    @op(cache=True, version="1.0")
    def cached_bar(message: str) -> str:
        print("bar was called")
        return f"message from bar: {message}"

    with lzy.workflow(workflow_name, interactive=False):
        for _ in range(3):
            assert bar_with_print(foo_with_print("number", 42)) == "message from bar: number is 42"
    with lzy.workflow(workflow_name, interactive=False):
        assert bar_with_print(foo_with_print("number", 42)) == "message from bar: number is 42"
    assert capsys.readouterr().out.count("foo was called") == 1

    before = len(lzy.runtime.executed)
    for _ in range(2):
        with lzy.workflow(workflow_name, interactive=False):
            assert cached_bar(foo_with_print("number", 42)) == "message from bar: number is 42"
    out = capsys.readouterr().out
    assert out.count("bar was called") == 1 and "foo was called" not in out
    assert lzy.runtime.executed[before:] == ["cached_bar"]


def test_scenario_two_executions_of_one_workflow_name(lzy, workflow_name):
    @op
    def return_42() -> int:
        return 42

    @op
    def return_13() -> int:
        return 13

    with lzy.workflow(workflow_name, interactive=False):
        assert return_42() == 42
    with lzy.workflow(workflow_name, interactive=False):
        assert return_13() == 13


def test_scenario_files_travel_by_content(lzy, workflow_name, tmp_path):
    """file_test/__init__.py:14-58 without the whiteboard half: a File nested in a dataclass argument, three File
    results, contents intact and independent of the caller's files"""
    import dataclasses

    from lzy_b200.types import File

    @dataclasses.dataclass
    class CustomFileField:
        file: File

    b_path, c_path = str(tmp_path / "b.txt"), str(tmp_path / "c.txt")

    @op
    def a(f1: CustomFileField) -> (File, File, File):
        with f1.file.open("a") as fl:
            fl.write("buzz")
        with open(b_path, "w") as fl:
            fl.write("Lol")
        with open(c_path, "w") as fl:
            fl.write("Kek")
        return f1.file, File(b_path), File(c_path)

    @op
    def b() -> (int, str):
        return 42, "42"

    src = tmp_path / "a.txt"
    src.write_text("fizz")
    with lzy.workflow(workflow_name, interactive=False):
        f1, f2, f3 = a(CustomFileField(File(str(src))))
        texts = []
        for f in (f1, f2, f3):
            with f.open("r") as fh:
                texts.append(fh.read())
        i, s = b()
        assert (int(i), str(s)) == (42, "42")
    assert texts == ["fizzbuzz", "Lol", "Kek"]
    assert src.read_text() == "fizz"  # the op appended to its own copy


def test_scenario_large_input_output_round_trip(lzy, workflow_name):
    """large_input_output/__init__.py:10-21 at a tenth of its size: a DataFrame goes in, comes back, equal"""
    import numpy as np
    import pandas as pd

    @op
    def accept_return(df: pd.DataFrame) -> pd.DataFrame:
        return df

    data = pd.DataFrame(np.random.default_rng(0).random((500_000, 3)), columns=["a", "b", "c"])
    with lzy.workflow(workflow_name, interactive=False):
        back = materialize(accept_return(data))
    assert back is not data and back.equals(data)


def test_scenario_custom_serializer_is_shipped_to_the_op_side_as_an_import(workflow_name):
    """custom_serializer/__init__.py:19-24 and registry.py:60-71: a serializer registered on
    `lzy.serializer_registry` travels in the request as (module, class, priority); the op side builds its own
    registry from those imports — here a fresh `LzySerializerRegistry` inside `startup.main` — and the payload on
    the wire is the custom serializer's bytes"""
    from lzy_b200 import _host as H
    from tests.support.custom_serializer import Point, PointSerializer
    from tests.support.worker_pool_runtime import WorkerPoolRuntime

    @op
    def mirror(p: Point) -> Point:
        return Point(p.y, p.x)

    rt = WorkerPoolRuntime(n_workers=1)
    try:
        lzy = Lzy(runtime=rt, storage_uri="mem://bucket/custom")
        lzy.serializer_registry.register_serializer(PointSerializer())
        assert [(i.module_name, i.class_name) for i in lzy.serializer_registry.imports()] == \
            [("tests.support.custom_serializer", "PointSerializer")]
        with lzy.workflow(workflow_name) as wf:
            out = mirror(Point(3, 4))
            assert materialize(out) == Point(4, 3)
            call = wf.owner.runtime.calls[-1]
            arg_uri = wf.snapshot.get(call.arg_entry_ids[0]).storage_uri
            ret_uri = wf.snapshot.get(call.entry_ids[0]).storage_uri
            assert H.blob_get(arg_uri) == b"point:3,4" and H.blob_get(ret_uri) == b"point:4,3"
    finally:
        rt.close()


@op
def _inner_double(x: int) -> int:
    return x * 2


@op
def _outer_with_its_own_workflow(x: int) -> int:
    rt = WorkerPoolRuntime(n_workers=1)
    try:
        with Lzy(runtime=rt, storage_uri="mem://bucket/inner").workflow("inner_" + uuid.uuid4().hex):
            return int(materialize(_inner_double(x))) + 1
    finally:
        rt.close()


def test_scenario_nested_workflow_inside_an_op(lzy, workflow_name):
    """nested_workflows/__init__.py: an op may open a workflow of its own (in the reference because it runs in its
    own process; here the workflow belongs to the op's thread) while the caller's stays active; outside an op a
    second simultaneous workflow is still refused (core/workflow.py:205-215)"""
    with lzy.workflow(workflow_name) as outer:
        assert materialize(_outer_with_its_own_workflow(20)) == 41
        assert type(outer).get_active() is outer
        with pytest.raises(RuntimeError, match="Simultaneous workflows are not supported"):
            with lzy.workflow(workflow_name + "_second"):
                pass


# ---- remaining cases of pylzy/tests/core/test_workflow.py that the shared scenarios do not already cover ----------
def test_tuple_with_ellipsis_unpacks_and_union_results_behave_like_their_values(lzy, workflow_name):
    from typing import List, Tuple, Union

    @op
    def returns_tuple() -> Tuple[List[int], ...]:
        return [1], [2]

    @op
    def union(p: bool) -> Union[str, int]:
        return "str" if p else 42

    with lzy.workflow(workflow_name):
        a, b = returns_tuple()
        s, n = union(True), union(False)
        assert sum(1 for _ in range(n)) == 42 and n == 42 and s == "str"
    assert a == [1] and b == [2]


def test_typed_dict_round_trip_and_refusal(lzy, workflow_name):
    from typing import Dict

    @op
    def echo(d: Dict[int, str]) -> Dict[int, str]:
        return d

    with lzy.workflow(workflow_name):
        res = echo(echo({1: "2"}))
    assert res == {1: "2"}
    with pytest.raises(TypeError):
        with lzy.workflow(workflow_name):
            echo({1: 1})


def test_proxies_inside_a_list_argument(lzy, workflow_name):
    """test_lazy_list (:435-447): results of other ops placed in a plain list argument"""
    from typing import List

    @op
    def foo() -> str:
        return "Foo:"

    @op
    def lens(strings: List[str], strings2: List[str]) -> (int, int):
        return len(strings), len(strings2)

    with lzy.workflow(workflow_name):
        a, b, c = foo(), foo(), foo()
        n1, n2 = lens([a, b, c], strings2=[a, b, c])
    assert (n1, n2) == (3, 3)


def test_returning_the_argument(lzy, workflow_name):
    @op
    def return_argument(arg: str) -> str:
        return arg

    with lzy.workflow(workflow_name):
        res = return_argument("str")
    assert res == "str"


# ---- the same stack with every worker behind its gRPC servers ---------------------------------------------------------
@pytest.fixture
def lzy_over_grpc():
    rt = WorkerPoolRuntime(n_workers=2, over_grpc=True)
    yield Lzy(runtime=rt, storage_uri="mem://bucket/grpc")
    rt.close()


def test_workflows_driven_through_worker_api_on_the_wire(lzy_over_grpc, workflow_name, capsys):
    """WorkerApi.Execute / LongRunningService.Get messages instead of method calls (what the reference's graph
    executor sends a worker VM): env vars ride in Operation.env.env, failures come back as rc + description and the
    caller still gets the op's own exception from its exception entry"""
    import os

    lzy = lzy_over_grpc

    @op
    def inc(x: int) -> int:
        print(os.environ.get("MODE", "-"))
        return x + 1

    @op
    def fail(x: int) -> int:
        raise KeyError("nope")

    with lzy.workflow(workflow_name).with_env_vars(MODE="wire"):
        assert materialize(inc(inc(inc(1)))) == 4
    assert capsys.readouterr().out.split() == ["wire", "wire", "wire"]
    with pytest.raises(KeyError, match="nope"):
        with lzy.workflow(workflow_name + "_f"):
            materialize(inc(fail(1)))
    assert lzy.runtime.executed.count("inc") == 3  # the consumer of the failed op never started


# ---- and with every part in its own process ------------------------------------------------------------------------------
@op
def _mp_double(x: int) -> int:
    return x * 2


@op
def _mp_join(a: str, b: str) -> str:
    return a + "|" + b


@op
def _mp_tag(t: str, suffix: str) -> str:
    return t + suffix


@op
def _mp_fail(x: int) -> int:
    raise LookupError(f"no {x}")


def test_workflows_on_a_channel_manager_process_and_worker_processes(tmp_path, workflow_name):
    """A channel-manager process, two worker processes, this process as the client: arguments and results under
    file:// URIs, consecutive tasks on different workers exchange data across the process boundary, an op's own
    exception comes back with its type, `cache=True` results are found again by a second workflow."""
    from lzy_b200.types import File
    from tests.support.worker_pool_runtime import ProcessPoolRuntime

    @op(cache=True, version="1")
    def cached_len(f: File) -> int:
        with f.open("rb") as fh:
            return len(fh.read())

    rt = ProcessPoolRuntime(str(tmp_path / "store"), n_workers=2)
    try:
        lzy = Lzy(runtime=rt, storage_uri=f"file://{tmp_path}/store")
        with lzy.workflow(workflow_name):
            y = _mp_double(_mp_double(_mp_double(5)))          # three hops, alternating workers
            s = _mp_tag("s", "0")
            d = _mp_join(_mp_tag(s, "1"), _mp_tag(s, "2"))     # a diamond
            assert materialize(y) == 40 and materialize(d) == "s01|s02"
        with pytest.raises(LookupError, match="no 7"):
            with lzy.workflow(workflow_name):
                materialize(_mp_double(_mp_fail(7)))
        src = tmp_path / "payload.bin"
        src.write_bytes(b"x" * 12345)
        for _ in range(2):
            with lzy.workflow(workflow_name):
                assert materialize(cached_len(File(str(src)))) == 12345
        assert rt.executed.count("cached_len") == 1 and rt.skipped_cached == 1
    finally:
        rt.close()


def test_arrays_travel_as_frame_streams_when_there_is_no_hbm(lzy, workflow_name):
    """ndarray / CPU tensor arguments and results through the serializers' byte-stream form (header + C-order bytes,
    the same layout a frame has in HBM): dtype, shape and bytes survive, non-contiguous inputs arrive contiguous"""
    import numpy as np
    import torch

    @op
    def scale(a: np.ndarray, t: torch.Tensor) -> (np.ndarray, torch.Tensor):
        assert a.flags["C_CONTIGUOUS"] and t.is_contiguous() and not t.is_cuda
        return a * 2, t + 1

    a = np.arange(3 * 5 * 7, dtype=np.int64).reshape(3, 5, 7)[:, ::2, :]      # a strided view
    t = torch.arange(24, dtype=torch.float32).reshape(4, 6).t()                # a transposed view
    with lzy.workflow(workflow_name):
        ra, rt = scale(a, t)
        ra, rt = materialize(ra), materialize(rt)
    assert ra.dtype == np.int64 and ra.shape == (3, 3, 7) and np.array_equal(ra, a * 2)
    assert rt.dtype == torch.float32 and tuple(rt.shape) == (6, 4) and torch.equal(rt, t + 1)


def test_identical_cached_calls_in_one_graph_run_once(lzy, workflow_name, capsys):
    """DataFlowGraph.java:51-72: a graph is keyed by data URI — the second of two equal cache=True calls submitted in
    one barrier is dropped, its consumers are wired to the first one's result."""
    @op(cache=True, version="1.0")
    def message_producer(name: str) -> str:
        print("message producer was called")
        return f"My name is {name}"

    @op(cache=False)
    def bar(message: str, name: str) -> str:
        return f"message from '{name}' bar: {message}"

    before = lzy.runtime.dropped_duplicates
    with lzy.workflow(workflow_name, interactive=False) as wf:
        first, second = message_producer("Graceful"), message_producer("Graceful")
        a, b = bar(first, "a"), bar(second, "b")
        wf.barrier()
        assert a == "message from 'a' bar: My name is Graceful" and b == "message from 'b' bar: My name is Graceful"
        assert second == "My name is Graceful" and first == second
    assert capsys.readouterr().out.count("message producer was called") == 1
    assert lzy.runtime.dropped_duplicates == before + 1
