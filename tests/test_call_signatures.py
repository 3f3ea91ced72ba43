"""Call-time signature inference and validation of ``@op`` calls, scenario for scenario with the reference's
pylzy/tests/core/test_calls.py:55-372 (implementation pylzy/lzy/api/v1/utils/types.py, signatures.py,
core/call.py:204-312): which names become entries, which types they are given (annotation, else inferred from
the value, else taken from the producing op for a proxy), and what is refused before anything runs.  A
recording runtime stands in (execute=False): only the queued calls are inspected."""
from typing import List, Optional, Tuple, Union

import pytest

from lzy_b200.api.v1 import Lzy, op
from lzy_b200.api.v1.utils.proxy_adapter import materialized
from tests.mocks import RuntimeMock


@op
def produce_int() -> int:
    return 1


@op
def takes_int(arg: int) -> str:
    pass


@op
def takes_float_and_rest(c: float, *args, **kwargs) -> None:
    pass


@op
def rest_only(*args) -> str:
    pass


@op
def unannotated(arg) -> str:
    pass


class Thing:
    pass


@op
def takes_thing(arg: Thing) -> Thing:
    pass


@pytest.fixture
def lzy():
    return Lzy(runtime=RuntimeMock(execute=False), storage_uri="mem://bucket/prefix")


def queued(lzy, i=0):
    return lzy.runtime.calls[i].signature.func


def test_missing_and_surplus_arguments_are_refused(lzy):
    with pytest.raises(KeyError, match="is required but not provided"):
        with lzy.workflow("test"):
            takes_int()
    with pytest.raises(KeyError, match="Unexpected argument"):
        with lzy.workflow("test"):
            takes_int(1, "s", 2)
    with pytest.raises(KeyError, match="Unexpected key argument"):
        with lzy.workflow("test"):
            rest_only(k=1, s="s", p=2)
    with pytest.raises(KeyError, match="Unexpected key argument"):
        with lzy.workflow("test"):
            takes_int(invalid_arg=1)


def test_defaults_travel_with_the_function_not_as_entries(lzy):
    @op
    def all_default(arg: int = 1) -> str:
        pass

    @op
    def one_default(arg: int, b: str = "default") -> str:
        pass

    with lzy.workflow("test"):
        all_default()
        one_default(1)
    f0, f1 = queued(lzy, 0), queued(lzy, 1)
    assert f0.output_types[0] is str and len(f0.input_types) == 0
    assert f1.output_types[0] is str and dict(f1.input_types) == {"arg": int}


def test_a_proxy_argument_takes_the_type_its_producer_declared(lzy):
    @op
    def produce_strs() -> List[str]:
        pass

    with lzy.workflow("test"):
        i = produce_int()
        assert materialized(i) is False
        takes_int(arg=i)
        v = produce_strs()
        assert materialized(v) is False
        unannotated(v)
    assert queued(lzy, 0).output_types[0] is int and len(queued(lzy, 0).input_types) == 0
    assert queued(lzy, 1).output_types[0] is str and dict(queued(lzy, 1).input_types) == {"arg": int}
    assert dict(queued(lzy, 3).input_types) == {"arg": List[str]}


@pytest.mark.parametrize("value, expected", [
    ([1, 2, 3], List[int]),
    ((1, 2, 3), Tuple[int, int, int]),
    (tuple(range(1000)), Tuple[int, ...]),
])
def test_types_of_unannotated_arguments_are_inferred_from_the_value(lzy, value, expected):
    with lzy.workflow("test"):
        unannotated(value)
    f = queued(lzy)
    assert f.output_types[0] is str and dict(f.input_types) == {"arg": expected}


def test_none_for_an_unannotated_argument(lzy):
    with lzy.workflow("test"):
        unannotated(None)
    assert issubclass(queued(lzy).input_types["arg"], type(None))


def test_generic_annotations_are_kept_and_checked_against_the_value(lzy):
    @op
    def strs(arg: List[str]) -> str:
        pass

    @op
    def ints(lst: List[int]) -> None:
        pass

    with lzy.workflow("test"):
        strs(["1", "2", "3"])
    assert dict(queued(lzy).input_types) == {"arg": List[str]}
    with pytest.raises(TypeError, match="Invalid types"):
        with lzy.workflow("test"):
            strs([1, 2, 3])
    with pytest.raises(TypeError):
        with lzy.workflow("test"):
            ints(["a", "b", "c"])


@pytest.mark.parametrize("annotation, values", [(Optional[str], (None, "str")), (Union[int, str], (1, "str"))])
def test_optional_and_union_annotations_stay_as_declared(lzy, annotation, values):
    @op
    def f(arg: annotation) -> annotation:
        pass

    with lzy.workflow("test"):
        for v in values:
            f(v)
    for i in range(len(values)):
        sig = queued(lzy, i)
        assert sig.output_types[0] == annotation and dict(sig.input_types) == {"arg": annotation}


def test_custom_classes(lzy):
    with lzy.workflow("test"):
        takes_thing(Thing())
    assert queued(lzy).output_types[0] is Thing and dict(queued(lzy).input_types) == {"arg": Thing}

    class Other:
        pass

    with pytest.raises(TypeError, match="Invalid types"):
        with lzy.workflow("test"):
            takes_thing(Other())
    with pytest.raises(TypeError, match="Invalid types"):  # a proxy is checked by the type it will have
        with lzy.workflow("test"):
            takes_thing(produce_int())


def test_varargs_and_kwargs_become_named_entries(lzy):
    with lzy.workflow("test"):
        takes_float_and_rest(2.0, *(1, 2, 3), **{"a": "a", "b": 2})
    f = queued(lzy)
    assert len(f.input_types) == 6 and f.input_types["a"] is str and f.input_types["b"] is int and f.input_types["c"] is float
    assert f.kwarg_names == ("a", "b")


def test_varargs_with_proxies(lzy):
    with lzy.workflow("test"):
        i = produce_int()
        takes_float_and_rest(2.0, *(i,), **{"a": i})
    f = queued(lzy, 1)
    assert len(f.input_types) == 3 and f.input_types["a"] is int and f.input_types["c"] is float and f.kwarg_names == ("a",)


def test_several_return_values_and_none(lzy):
    @op
    def pair() -> (int, str):
        pass

    @op
    def nothing() -> None:
        pass

    with lzy.workflow("test"):
        pair()
        assert nothing() is None
    assert tuple(queued(lzy).output_types) == (int, str)


def test_a_failing_op_is_still_only_queued(lzy):
    @op
    def raises() -> int:
        raise RuntimeError("Bad exception")

    with lzy.workflow("test"):
        raises()
    assert queued(lzy).output_types[0] is int


def test_outside_a_workflow_and_bad_declarations(lzy):
    assert produce_int() == 1
    with pytest.raises(TypeError, match=r"Return type is not annotated for function `no_hint` at .*"):
        @op
        def no_hint():
            pass
    with pytest.raises(ValueError, match="Invalid workflow name. Name can contain only"):
        with lzy.workflow("test test"):
            pass


def test_annotation_helpers_answer_as_the_reference_module_does():
    """tests/golden/reference_types_vectors.json holds what pylzy/lzy/api/v1/utils/types.py itself (executed in the
    authoring container, tests/golden/make_golden.py) answers to tests/golden/types_battery.py: real types of unions
    and generics, declared return types, defaults, the subtype relation with the empty-container rule, and whether
    the serializer picked for an annotation takes a runtime type"""
    import importlib.util
    import json
    from pathlib import Path

    from lzy_b200.api.v1.utils import types as T
    from lzy_b200.proxy_result import Absence, Result
    from lzy_b200.serialization.types import EmptyContent

    golden = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("types_battery", golden / "types_battery.py")
    battery = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(battery)
    ours = battery.run(dict(
        infer_real_types=T.infer_real_types, infer_return_type=T.infer_return_type, get_default_args=T.get_default_args,
        is_subtype=T.is_subtype, check_types_serialization_compatible=T.check_types_serialization_compatible,
        EmptyContent=EmptyContent, Absence=Absence, Result=Result))
    reference = json.loads((golden / "reference_types_vectors.json").read_text())
    assert json.loads(json.dumps(ours)) == reference


def test_argument_binding_answers_as_the_reference_function_does():
    """tests/golden/reference_call_vectors.json: 55 calls bound by pylzy/lzy/core/call.py's own
    infer_and_validate_call_signature (executed in the authoring container) — refusals with their exact messages,
    the names arguments get as entries (made-up names of varargs normalised), the types recorded for them, the values
    that travel as local data (proxies stay proxies, materialised ones become values)"""
    import importlib.util
    import json
    from pathlib import Path

    from lzy_b200.api.v1.utils import proxy_adapter
    from lzy_b200.core.call import infer_and_validate_call_signature
    from lzy_b200.proxy_result import Result
    from lzy_b200.serialization.registry import LzySerializerRegistry

    golden = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("call_battery", golden / "call_battery.py")
    battery = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(battery)
    ours = battery.run(dict(bind=infer_and_validate_call_signature, lzy_proxy=proxy_adapter.lzy_proxy,
                            materialize=proxy_adapter.materialize, is_lzy_proxy=proxy_adapter.is_lzy_proxy,
                            Result=Result, registry=LzySerializerRegistry()))
    reference = json.loads((golden / "reference_call_vectors.json").read_text())
    assert set(ours) == set(reference) and len(reference) == 55
    for label, expected in reference.items():
        assert json.loads(json.dumps(ours[label])) == expected, label


def test_recording_a_call_touches_the_snapshot_as_the_reference_class_does():
    """tests/golden/reference_lzycall_vectors.json: pylzy/lzy/core/call.py's own LzyCall constructed (in the authoring
    container) over a recording snapshot — entries are created for local arguments, results and the exception in
    that order and under those names, local arguments are written at construction, arguments that are results of
    other ops reuse their entries, the environment is combined Lzy → workflow → call and validated"""
    import importlib.util
    import json
    from pathlib import Path

    from lzy_b200.api.v1.utils import proxy_adapter
    from lzy_b200.core import call as core_call
    from lzy_b200.proxy_result import Result
    from lzy_b200.serialization.registry import LzySerializerRegistry

    golden = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("lzycall_battery", golden / "lzycall_battery.py")
    battery = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(battery)
    ours = battery.run(dict(bind=core_call.infer_and_validate_call_signature, LzyCall=core_call.LzyCall,
                            lzy_proxy=proxy_adapter.lzy_proxy, Result=Result, registry=LzySerializerRegistry()))
    reference = json.loads((golden / "reference_lzycall_vectors.json").read_text())
    assert json.loads(json.dumps(ours)) == reference
