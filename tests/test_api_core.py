"""Python API semantics that do not need a GPU, on a recording/in-process
RuntimeMock like the reference's pylzy/tests/core/test_calls.py and pylzy/tests/api/v1/test_entries.py
(URI rules :91-330, call-time snapshot of arguments :331-384)."""
import hashlib
import os
import uuid
from typing import Dict, List, Optional, Tuple, Union

import pytest

from lzy_b200.api.v1 import Lzy, LzyWorkflow, materialize, materialized, op
from lzy_b200.api.v1.utils.proxy_adapter import get_proxy_entry_id, is_lzy_proxy
from tests.mocks import RuntimeMock


@op
def foo(name: str, param: int) -> str:
    return f"{name}: {param}"


@op
def foo_varargs(*args) -> str:
    return ", ".join(str(a) for a in args)


@op
def accept_arr(ints: List[int]) -> List[int]:
    return ints


@op
def accept_arr_and_modify(ints: List[int]) -> List[int]:
    ints.append(len(ints))
    return ints


@pytest.fixture
def lzy():
    return Lzy(runtime=RuntimeMock(), storage_uri="mem://bucket/prefix")


def wf_name():
    return f"wf_{uuid.uuid4().hex}"


def test_return_annotation_is_mandatory():
    with pytest.raises(TypeError, match="Return type is not annotated"):
        @op
        def no_annotation():
            pass


def test_op_outside_workflow_is_a_plain_call():
    assert foo("a", 1) == "a: 1"
    assert foo.__name__ == "foo"


def test_literal_tuple_annotation_means_several_outputs(lzy):
    @op
    def two() -> (str, int):
        return "s", 42

    @op
    def one() -> Tuple[str, int]:
        return "s", 42

    with lzy.workflow(wf_name()) as wf:
        a, b = two()
        t = one()
        assert len(wf.call_queue[0].entry_ids) == 2 and len(wf.call_queue[1].entry_ids) == 1
    assert (a, b) == ("s", 42)
    x, y = t  # one output that unpacks by iteration
    assert (x, y) == ("s", 42)


def test_none_annotation_returns_real_none_but_is_queued(lzy):
    @op
    def nothing() -> None:
        pass

    with lzy.workflow(wf_name()) as wf:
        r = nothing()
        assert r is None and len(wf.call_queue) == 1
    assert len(lzy.runtime.calls) == 1


def test_lazy_until_touched_and_eager(lzy):
    with lzy.workflow(wf_name()) as wf:
        s = foo("x", 1)
        assert is_lzy_proxy(s) and not materialized(s) and len(wf.call_queue) == 1
        assert lzy.runtime.calls == []
        assert s == "x: 1"  # touching materialises: barrier, then get_data
        assert materialized(s) and len(wf.call_queue) == 0
    eager = Lzy(runtime=RuntimeMock(), storage_uri="mem://b")
    with eager.workflow(wf_name(), eager=True) as wf:
        s = foo("y", 2)
        assert type(s) is str and s == "y: 2" and len(wf.call_queue) == 0


def test_proxy_args_are_not_put_again_and_materialized_proxies_are_local_data(lzy):
    with lzy.workflow(wf_name()) as wf:
        a = accept_arr([1, 2, 3])
        b = accept_arr(a)
        c0, c1 = wf.call_queue
        assert c1.arg_entry_ids[0] == c0.entry_ids[0] == get_proxy_entry_id(a)
        a.append(4)  # materialises `a`, then mutates the local object
        c = accept_arr(a)
        assert wf.call_queue[-1].arg_entry_ids[0] != get_proxy_entry_id(a)
    assert materialize(c) == [1, 2, 3, 4] and materialize(b) == [1, 2, 3]


def test_argument_type_validation_at_call_time(lzy):
    with lzy.workflow(wf_name()):
        with pytest.raises(TypeError, match="Invalid types: argument param"):
            foo("a", "not an int")
        with pytest.raises(KeyError, match="required but not provided"):
            foo("a")
        with pytest.raises(KeyError, match="Unexpected argument"):
            foo("a", 1, 2)
        with pytest.raises(KeyError, match="Unexpected key argument"):
            foo("a", 1, nope=3)
        with pytest.raises(TypeError):
            accept_arr(["str"])
        accept_arr([])  # List[EmptyContent] fits List[int]


def test_optional_union_dict_annotations(lzy):
    @op
    def opt(a: Optional[str]) -> str:
        return "none" if a is None else "not none"

    @op
    def uni(p: Union[str, int]) -> bool:
        return isinstance(p, str)

    @op
    def d(x: Dict[int, str]) -> Dict[int, str]:
        return x

    with lzy.workflow(wf_name()):
        assert opt(None) == "none" and opt("x") == "not none"
        assert uni("s") == True and uni(1) == False  # noqa: E712
        assert d({1: "2"}) == {1: "2"} and d({}) == {}
        with pytest.raises(TypeError):
            d({1: 1})


def test_same_local_data_stored_once_and_uri_is_prefix_plus_hash(lzy):
    with lzy.workflow("test") as wf:
        foo("hehe", 42)
        foo_varargs(42)
        q = list(wf.call_queue)
    snap = wf.snapshot
    uri_1 = snap.get(q[0].arg_entry_ids[1]).storage_uri
    uri_2 = snap.get(q[1].arg_entry_ids[0]).storage_uri
    assert uri_1 == uri_2 and lzy.runtime.store_counts[uri_1] == 1
    assert uri_1.startswith("mem://bucket/prefix/lzy_runs/test/inputs/")
    assert snap.get(q[0].arg_entry_ids[1]).data_hash == uri_1.rsplit("/", 1)[1]


def test_result_uris_without_cache(lzy):
    with lzy.workflow("test") as wf:
        foo("a", 1)
        foo("a", 1)
        calls = list(wf.call_queue)
    snap = wf.snapshot
    uris = [snap.get(c.entry_ids[0]).storage_uri for c in calls]
    assert uris[0] != uris[1]
    for c, u in zip(calls, uris):
        assert u == f"mem://bucket/prefix/lzy_runs/test/ops/foo/{c.id}/return_0"
        assert snap.get(c.entry_ids[0]).data_hash == hashlib.md5(u.encode()).hexdigest()
        assert snap.get(c.exception_id).storage_uri == f"mem://bucket/prefix/lzy_runs/test/ops/foo/{c.id}/exception"


def test_result_uris_with_cache(lzy, monkeypatch):
    @op(cache=True, version="1.7")
    def cached(a: int, *, k: str = "z", b: int = 0) -> int:
        return a + b

    with lzy.workflow("test") as wf:
        cached(1, k="kk", b=2)
        cached(1, b=2, k="kk")
        cached(2, k="kk", b=2)
        calls = list(wf.call_queue)
    snap = wf.snapshot
    uris = [snap.get(c.entry_ids[0]).storage_uri for c in calls]
    assert uris[0] == uris[1] != uris[2]
    h = lambda c, e: snap.get(e).data_hash
    c = calls[0]
    expect = "_".join([h(c, c.arg_entry_ids[0]), f"b:{h(c, c.kwarg_entry_ids['b'])}", f"k:{h(c, c.kwarg_entry_ids['k'])}"])
    assert uris[0] == f"mem://bucket/prefix/lzy_runs/test/ops/cached/1.7/{expect}/return_0"
    monkeypatch.setenv("LZY_USER", "igor")
    with Lzy(runtime=RuntimeMock(), storage_uri="mem://bucket/prefix").workflow("test") as wf2:
        cached(1, k="kk", b=2)
        c2 = wf2.call_queue[0]
    assert wf2.snapshot.get(c2.entry_ids[0]).storage_uri.startswith("mem://bucket/prefix/igor/lzy_runs/test/ops/cached/1.7/")


def test_local_data_changed_by_ref_is_snapshotted_at_call_time(lzy):
    with lzy.workflow("wf") as wf:
        ints = []
        for i in range(5):
            ints.append(i)
            accept_arr(ints)
        calls = list(wf.call_queue)
    arg_uris = [wf.snapshot.get(c.arg_entry_ids[0]).storage_uri for c in calls]
    assert len(set(arg_uris)) == 5
    assert [materialize_result(lzy, wf, c) for c in calls] == [[0], [0, 1], [0, 1, 2], [0, 1, 2, 3], [0, 1, 2, 3, 4]]


def materialize_result(lzy, wf, call):
    import cloudpickle
    return cloudpickle.loads(lzy.runtime.blobs[wf.snapshot.get(call.entry_ids[0]).storage_uri])


def test_local_data_can_not_be_changed_by_ref_inside_op(lzy):
    with lzy.workflow("wf") as wf:
        ints = []
        for i in range(5):
            accept_arr_and_modify(ints)
        calls = list(wf.call_queue)
    assert ints == []  # the op mutated its own copy
    assert len({wf.snapshot.get(c.arg_entry_ids[0]).storage_uri for c in calls}) == 1
    assert all(materialize_result(lzy, wf, c) == [0] for c in calls)


def test_only_one_active_workflow(lzy):
    with lzy.workflow(wf_name()):
        with pytest.raises(RuntimeError, match="Simultaneous workflows are not supported"):
            with lzy.workflow(wf_name()):
                pass
    assert LzyWorkflow.get_active() is None
    with lzy.workflow(wf_name()):  # the failed inner enter did not poison the state
        pass


def test_a_refused_second_workflow_leaves_the_active_one_running(lzy):
    with lzy.workflow(wf_name()):
        with pytest.raises(RuntimeError, match="Simultaneous workflows are not supported"):
            with lzy.workflow(wf_name()):
                pass
        assert lzy.runtime.aborted == 0  # the refused workflow started nothing, so it had nothing to abort
        assert foo("a", 1) == "a: 1"    # and the first one still runs ops
    assert lzy.runtime.aborted == 0 and lzy.runtime.finished == 1


def test_error_on_start_hands_the_decision_to_the_runtime(lzy, monkeypatch):
    """pylzy/tests/api/v1/test_remote_runtime.py:65-76: the body is not entered; the workflow tells the runtime to abort
    (as the reference's does — what an abort without a start means is the runtime's business: RemoteRuntime and
    B200Runtime do nothing then), and no workflow is left active"""
    async def refuse(workflow):
        raise ConnectionError("cannot start")

    monkeypatch.setattr(lzy.runtime, "start", refuse)
    entered = False
    with pytest.raises(ConnectionError):
        with lzy.workflow(wf_name()):
            entered = True
    assert not entered and lzy.runtime.aborted == 1 and lzy.runtime.finished == 0 and LzyWorkflow.get_active() is None


def test_invalid_workflow_name(lzy):
    with pytest.raises(ValueError, match="Invalid workflow name"):
        lzy.workflow("bad name!")


def test_exception_in_with_block_aborts(lzy):
    with pytest.raises(ZeroDivisionError):
        with lzy.workflow(wf_name()):
            foo("a", 1)
            1 / 0
    assert lzy.runtime.aborted == 1 and lzy.runtime.finished == 0 and lzy.runtime.calls == []
    assert LzyWorkflow.get_active() is None


def test_nested_op_call_inside_op_is_plain(lzy):
    @op
    def outer(x: int) -> str:
        return foo("inner", x)  # no workflow is active inside an op body

    with lzy.workflow(wf_name()):
        r = outer(5)
    assert r == "inner: 5" and len(lzy.runtime.calls) == 1


def test_serializer_registry_rules():
    from lzy_b200.serialization import LzySerializerRegistry
    from lzy_b200.serialization.api import Serializer
    from tests.support.custom_serializer import Point, PointSerializer

    reg = LzySerializerRegistry()

    class InMain(PointSerializer):
        pass

    InMain.__module__ = "__main__"
    with pytest.raises(ValueError, match="__main__"):
        reg.register_serializer(InMain())
    reg.register_serializer(PointSerializer())
    assert isinstance(reg.find_serializer_by_type(Point), PointSerializer)
    imps = reg.imports()
    assert [(i.module_name, i.class_name) for i in imps] == [("tests.support.custom_serializer", "PointSerializer")]
    other = LzySerializerRegistry()
    other.load_imports(imps)
    assert isinstance(other.find_serializer_by_type(Point), PointSerializer)
    # nested values of a registered type go through their serializer inside pickles (copyreg hook)
    import cloudpickle
    blob = cloudpickle.dumps({"p": Point(1, 2)})
    assert b"point:1,2" in blob and cloudpickle.loads(blob)["p"] == Point(1, 2)
    reg.unregister_serializer(PointSerializer())
    other.unregister_serializer(PointSerializer())
    assert reg.imports() == []


def test_install_as_lzy_aliases_the_package():
    """unmodified user code imports `lzy.api.v1` (pylzy/readme.md quick start)"""
    import subprocess
    import sys
    code = ("import lzy_b200; lzy_b200.install_as_lzy();"
            "from lzy.api.v1 import Lzy, op;"
            "from lzy.types import File;"
            "import lzy_b200.api.v1 as v1; assert Lzy is v1.Lzy and op is v1.op; print('ok')")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(__import__("pathlib").Path(__file__).resolve().parent.parent))
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr


def test_runtime_config_env_overrides():
    """LZY_B200_* environment overrides for deployment knobs; explicit arguments win"""
    from lzy_b200.api.v1.b200.runtime import B200RuntimeConfig
    env = {"LZY_B200_GPUS": "2, 3", "LZY_B200_HBM_BUDGET_BYTES": "1.5g", "LZY_B200_SPILL_DIR": "/nvme/spill",
           "LZY_B200_PREFETCH": "off", "LZY_B200_COPY_ENGINE": "dma", "LZY_B200_FANOUT_THRESHOLD": "5", "UNRELATED": "x"}
    c = B200RuntimeConfig.from_env(env)
    got = (c.gpus, c.hbm_budget_bytes, c.spill_dir, c.prefetch, c.copy_engine, c.fanout_threshold)
    assert got == ([2, 3], int(1.5 * (1 << 30)), "/nvme/spill", False, "dma", 5)
    assert c.placement == "round_robin" and c.chunk_bytes == 0  # untouched fields keep their defaults
    assert B200RuntimeConfig.from_env(env, gpus=[0], prefetch=True).gpus == [0]
    assert B200RuntimeConfig.from_env(env, prefetch=True).prefetch is True
    assert B200RuntimeConfig.from_env({"LZY_B200_GPUS": "all"}).gpus is None
    assert B200RuntimeConfig.from_env({}) == B200RuntimeConfig(user_id=B200RuntimeConfig.from_env({}).user_id)
    for bad in ({"LZY_B200_PLACEMENT": "random"}, {"LZY_B200_PREFETCH": "maybe"}, {"LZY_B200_CHUNK_BYTES": "2x"}):
        with pytest.raises(ValueError, match="LZY_B200_"):
            B200RuntimeConfig.from_env(bad)


def test_env_shortcut_decorators_and_their_refusals():
    """pylzy/lzy/env/shortcuts.py: `@env_vars(...)` / `@provisioning(...)` above `@op` (or applied to Lzy / a
    workflow); anything else is refused"""
    from lzy_b200.api.v1 import Provisioning, env_vars, provisioning

    @env_vars(A="1")
    @provisioning(gpu_count=2)
    @op
    def f() -> int:
        return 1

    assert f.env.get_env_vars() == {"A": "1"} and f.env.get_provisioning().gpu_count == 2 and f() == 1
    lzy = env_vars({"B": "2"})(Lzy(runtime=RuntimeMock()))
    assert lzy.env.get_env_vars() == {"B": "2"}
    wf = provisioning(Provisioning(gpu_index=3))(lzy.workflow(wf_name()))
    assert wf.env.get_provisioning().gpu_index == 3
    with pytest.raises(TypeError, match="can be applied only to"):
        env_vars(A="1")(lambda: None)
    with pytest.raises(TypeError):
        f.with_env_vars({"A": 1})
    with pytest.raises(NotImplementedError):
        f.with_env_vars({"A": "1"}, B="2")


def test_applied_env_shares_agreeing_values_and_serialises_conflicting_ones():
    """ops of one worker share a process: a variable belongs to the ops running with it"""
    import threading
    import time

    from lzy_b200.env import applied_env
    assert "LZYB_T" not in os.environ
    with applied_env(None):
        assert "LZYB_T" not in os.environ
    seen, order, gate = [], [], threading.Event()

    def op_body(value, hold):
        with applied_env({"LZYB_T": value}):
            order.append(("in", value))
            seen.append(os.environ["LZYB_T"])
            if hold:
                gate.wait(5)
            order.append(("out", value))

    a = threading.Thread(target=op_body, args=("1", True))
    a.start()
    while ("in", "1") not in order:
        time.sleep(0.001)
    b = threading.Thread(target=op_body, args=("1", False))  # same value: shares it, does not wait
    b.start()
    b.join(5)
    assert not b.is_alive() and os.environ["LZYB_T"] == "1"     # still set: `a` is using it
    c = threading.Thread(target=op_body, args=("2", False))  # other value: waits for `a`
    c.start()
    time.sleep(0.05)
    assert ("in", "2") not in order
    gate.set()
    a.join(5)
    c.join(5)
    last_out_1 = max(i for i, e in enumerate(order) if e == ("out", "1"))
    assert seen == ["1", "1", "2"] and last_out_1 < order.index(("in", "2"))
    assert "LZYB_T" not in os.environ
    os.environ["LZYB_T"] = "before"
    try:
        with applied_env({"LZYB_T": "during"}):
            assert os.environ["LZYB_T"] == "during"
        assert os.environ["LZYB_T"] == "before"
    finally:
        del os.environ["LZYB_T"]


def test_workflow_declarations_are_frozen_once_it_is_in_use(lzy):
    wf = lzy.workflow(wf_name()).with_env_vars(A="1")  # before use: a new object each time
    assert wf.env.get_env_vars() == {"A": "1"} and wf.with_env_vars(B="2") is not wf
    with wf:
        with pytest.raises(RuntimeError, match="forbidden to change LzyWorkflow fields"):
            wf.with_env_vars(C="3")


def test_dataflow_plan_is_keyed_by_uri():
    """lzy-service DataFlowGraph.java:51-72: one supplier per output URI; an op whose outputs are all supplied is
    dropped (its results point at the supplier); a partial overlap is an error; cached results are not resubmitted."""
    from types import SimpleNamespace
    from lzy_b200.core.dataflow import plan_dataflow

    uri = {"a0": "s/f/h1/return_0", "b0": "s/f/h1/return_0", "c0": "s/g/1/return_0", "d0": "s/f/h2/return_0",
           "e0": "s/p/h/return_0", "e1": "s/p/h/return_1", "f0": "s/p/h/return_0", "f1": "s/p/h/return_1",
           "g0": "s/old/return_0"}

    def call(cache, *eids):
        return SimpleNamespace(cache=cache, entry_ids=list(eids))

    a, b, c, d = call(True, "a0"), call(True, "b0"), call(False, "c0"), call(True, "d0")
    e, f, g, none = call(True, "e0", "e1"), call(True, "f0", "f1"), call(True, "g0"), call(False)
    plan = plan_dataflow([a, b, c, d, e, g, f, none, none], uri.__getitem__, {"s/old/return_0"}.__contains__)
    assert plan.calls == [a, c, d, e, none, none] and plan.duplicates == [b, f] and plan.cached == [g]
    assert plan.supplier == {"a0": 0, "b0": 0, "c0": 1, "d0": 2, "e0": 3, "e1": 3, "f0": 3, "f1": 3}
    assert plan.outputs_of(0) == ["a0", "b0"] and plan.outputs_of(3) == ["e0", "e1", "f0", "f1"] and plan.outputs_of(1) == ["c0"]
    assert plan.dropped == [g, b, f]

    # an op with results but cache=False whose URI is in storage still runs (only cache=True consults storage)
    assert plan_dataflow([call(False, "g0")], uri.__getitem__, lambda u: True).calls != []
    # supplying one of two URIs that are already supplied is refused
    uri["h0"], uri["h1"] = "s/p/h/return_0", "s/other/return_1"
    with pytest.raises(RuntimeError, match="Output slot with uri 's/p/h/return_0' already exists"):
        plan_dataflow([e, call(True, "h0", "h1")], uri.__getitem__, lambda u: False)


def test_hbm_runtime_lifecycle_without_a_device(monkeypatch):
    """pylzy/lzy/api/v1/remote/runtime.py:141-147,204-243 (and tests/api/v1/test_remote_runtime.py:65-76): exec on a
    runtime that is not running is refused, abort / finish of a runtime that never started do nothing, each workflow
    has its own execution — also when an op opens a nested one on the same runtime."""
    import asyncio
    from types import SimpleNamespace
    from lzy_b200.api.v1.b200.runtime import B200Runtime

    rt = B200Runtime()
    monkeypatch.setattr(rt, "_ensure_ready", lambda: None)
    destroyed = []
    rt.cm = SimpleNamespace(destroy_all=lambda execution_id, reason: destroyed.append((execution_id, reason)))
    run = asyncio.run

    run(rt.abort())
    run(rt.finish())
    assert destroyed == []
    with pytest.raises(ValueError, match="Runtime is not running"):
        run(rt.exec([], lambda step: None))

    outer, inner = SimpleNamespace(name="outer"), SimpleNamespace(name="inner")
    outer_id = run(rt.start(outer))
    inner_id = run(rt.start(inner))
    assert outer_id != inner_id and rt._execution_of(outer) == outer_id and rt._execution_of(inner) == inner_id
    run(rt.finish())
    assert destroyed == [(inner_id, "finished")] and rt._execution_of(outer) == outer_id
    with pytest.raises(ValueError, match="Runtime is not running"):
        rt._execution_of(inner)
    run(rt.abort())
    assert destroyed == [(inner_id, "finished"), (outer_id, "aborted")]
    run(rt.abort())  # a second abort finds nothing to do
    assert len(destroyed) == 2

    # told whose end it is (LzyWorkflow does that when the runtime offers it), the runtime ends that one only
    outer_id, inner_id = run(rt.start(outer)), run(rt.start(inner))
    run(rt.abort_workflow(SimpleNamespace(name="never started")))
    assert len(destroyed) == 2 and rt._execution_of(inner) == inner_id
    run(rt.finish_workflow(outer))  # not the innermost
    assert destroyed[-1] == (outer_id, "finished") and rt._execution_of(inner) == inner_id
    run(rt.abort_workflow(inner))
    assert destroyed[-1] == (inner_id, "aborted")
    with pytest.raises(ValueError, match="Runtime is not running"):
        run(rt.exec([], lambda step: None))


def test_result_uris_are_the_ones_the_reference_workflow_assigns(monkeypatch):
    """tests/golden/reference_uri_vectors.json: what pylzy/lzy/core/workflow.py's own URI rules (executed in the
    authoring container) give the results and exceptions of twelve calls on two storages, without and with
    LZY_USER — cached and uncached, argument order, keyword order, versions, no arguments at all"""
    import importlib.util
    import json
    from pathlib import Path
    from types import SimpleNamespace

    golden = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("uri_battery", golden / "uri_battery.py")
    battery = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(battery)

    def make(name, storage, snapshot):
        wf = LzyWorkflow(name=name, owner=SimpleNamespace(storage_uri=storage), env=None)
        wf._snapshot = snapshot
        return wf

    for expected in json.loads((golden / "reference_uri_vectors.json").read_text()):
        monkeypatch.delenv("LZY_USER", raising=False)
        if expected["user"] is not None:
            monkeypatch.setenv("LZY_USER", expected["user"])
        ours = battery.run(dict(make_workflow=make, assign=lambda wf, call: wf._assign_uris(call)), expected["user"])
        assert ours == expected


def test_decorating_and_calling_an_op_behaves_as_the_reference_wrapper_does():
    """tests/golden/reference_wrapper_vectors.json: pylzy's own ``op`` and ``LazyCallWrapper`` (executed in the authoring
    container) on tests/golden/wrapper_battery.py — recorded result types and options, name and docstring carried over,
    the refusal of an unannotated function, plain calls outside a workflow; inside one every call is registered once
    and answered with one proxy per result bound to the call's result entries (a tuple for several), the real None
    for ``-> None``, values instead of proxies when the workflow is eager"""
    import importlib.util
    import json
    from pathlib import Path

    from lzy_b200.api.v1.utils import proxy_adapter
    from lzy_b200.proxy_result import Result
    from lzy_b200.serialization.registry import LzySerializerRegistry

    class Env:
        def combine(self, other):
            return self

        def validate(self):
            pass

        def get_env_vars(self):
            return {}

    golden = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("wrapper_battery", golden / "wrapper_battery.py")
    battery = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(battery)
    assert LzyWorkflow.instance is None
    ours = battery.run(dict(op=op, LzyWorkflow=LzyWorkflow, is_lzy_proxy=proxy_adapter.is_lzy_proxy,
                            get_proxy_entry_id=proxy_adapter.get_proxy_entry_id, Result=Result, Env=Env,
                            registry=LzySerializerRegistry()))
    assert LzyWorkflow.instance is None
    assert json.loads(json.dumps(ours)) == json.loads((golden / "reference_wrapper_vectors.json").read_text())


def test_workflow_lifecycle_drives_a_runtime_as_the_reference_workflow_does(monkeypatch):
    """tests/golden/reference_lifecycle_vectors.json: pylzy's own LzyWorkflow (executed in the authoring container) over a
    recording runtime — start on entry; queued calls run on exit, on barrier() (an empty queue asks nobody) or one by
    one when eager; finish after a clean block; abort — without running the queue — after an exception in the block,
    after a failed graph, and after a start that raised; nothing left active; local data under
    ``<storage>/[<user>/]lzy_runs/<name>/inputs/``"""
    import importlib.util
    import json
    from pathlib import Path
    from types import SimpleNamespace

    golden = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("lifecycle_battery", golden / "lifecycle_battery.py")
    battery = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(battery)

    def make(name, runtime, log, eager):
        snapshot = SimpleNamespace(prefix=None, entries={})
        snapshot.get = snapshot.entries.__getitem__

        def make_snapshot(workflow, prefix):
            snapshot.prefix = prefix
            return snapshot

        owner = SimpleNamespace(storage_uri="s3://bucket/prefix", runtime=runtime, make_snapshot=make_snapshot)
        return LzyWorkflow(name=name, owner=owner, env=None, eager=eager)

    for expected in json.loads((golden / "reference_lifecycle_vectors.json").read_text()):
        monkeypatch.delenv("LZY_USER", raising=False)
        if expected["user"] is not None:
            monkeypatch.setenv("LZY_USER", expected["user"])
        ours = battery.run(dict(make_workflow=make, active=lambda: getattr(LzyWorkflow.get_active(), "name", None)),
                           expected["user"])
        assert json.loads(json.dumps(ours)) == expected


def test_user_serializers_are_tracked_and_shipped_as_the_reference_registry_does():
    """tests/golden/reference_registry_vectors.json: pylzy's own LzySerializerRegistry (executed in the authoring container
    over a stand-in for serialzy's base registry) — built-ins are not shipped, user serializers are, as (module, class,
    priority); load_imports brings them up elsewhere; ``__main__`` is refused with the reference's message;
    unregistering forgets and removes the copyreg hook; a registered type nested in a pickle round-trips"""
    import importlib.util
    import json
    from pathlib import Path

    from lzy_b200.serialization.registry import LzySerializerRegistry

    golden = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("registry_battery", golden / "registry_battery.py")
    battery = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(battery)
    ours = battery.run(dict(new_registry=LzySerializerRegistry))
    assert json.loads(json.dumps(ours)) == json.loads((golden / "reference_registry_vectors.json").read_text())
