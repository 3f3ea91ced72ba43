"""One battery of ``@op`` calls, bound twice: by tests/golden/make_golden.py with the *reference's own*
``infer_and_validate_call_signature`` (pylzy/lzy/core/call.py:253-313, loaded by file path in the authoring container
— its environment / logging / snapshot imports stubbed, `serialzy.types.get_type` and `EmptyContent` supplied by this
package because serialzy itself is absent) to produce tests/golden/reference_call_vectors.json, and by
tests/test_call_signatures.py with lzy_b200.core.call's, which must answer the same: which calls are refused
(KeyError / TypeError, message included), under which names the arguments become entries, with which types, and
which values travel as local data.

`ns`: ``bind`` (the function under test), ``lzy_proxy`` / ``materialize`` (that side's proxy adapter), ``registry``."""
from types import SimpleNamespace
from typing import Any, Dict, List, Optional, Union


class Base:
    def __repr__(self):
        return type(self).__name__ + "()"


class Derived(Base):
    pass


def f0() -> int: ...
def f1(a: int) -> int: ...
def f2(a: int, b: str = "x") -> int: ...
def f3(a, b) -> int: ...
def f4(*args) -> int: ...
def f5(a: int, *rest) -> int: ...
def f6(a: int, **named) -> int: ...
def f7(a: List[int]) -> int: ...
def f8(a: Optional[int]) -> int: ...
def f9(a: Base) -> int: ...
def f10(a: Union[int, str], *, k: int = 0) -> int: ...
def f11(a: int, b: int = 1, *rest, k: str, **named) -> int: ...


def _type_name(t) -> str:
    return getattr(t, "__name__", None) if isinstance(t, type) else repr(t).replace("typing.", "")


def run(ns: Dict[str, Any]) -> Dict[str, Any]:
    entries = {"e-int": int, "e-str": str, "e-list": List[int], "e-base": Base}
    snapshot = SimpleNamespace(get=lambda eid: SimpleNamespace(typ=entries[eid]))
    # the proxy adapter only touches wflow when a proxy is materialised, and these carry their values
    values = {"e-int": 7, "e-str": "seven", "e-list": [7], "e-base": Derived()}

    def px(eid, touched=False):
        from lzy_b200.api.v1.utils.types import infer_real_types  # same answers as the reference's (types battery)
        p = ns["lzy_proxy"](eid, infer_real_types(entries[eid]), None, ns["Result"](values[eid]))
        if touched:
            ns["materialize"](p)
        return p

    def show(v):
        return "<proxy>" if ns["is_lzy_proxy"](v) else repr(v)

    def bind(f, *args, **kwargs):
        try:
            sig = ns["bind"](f, *args, output_type=(int,), snapshot=snapshot, serializer_registry=ns["registry"], **kwargs)
        except (KeyError, TypeError) as e:
            return {"refused": f"{type(e).__name__}: {e.args[0]}".replace("typing.", "")}
        n_declared = f.__code__.co_argcount
        rename = {name: f"*{i}" for i, name in enumerate(sig.func.arg_names[n_declared:])}  # made-up vararg names
        fix = lambda n: rename.get(n, n)  # noqa: E731
        return {"arg_names": [fix(n) for n in sig.func.arg_names], "kwarg_names": list(sig.func.kwarg_names),
                "input_types": {fix(n): _type_name(t) for n, t in sig.func.input_types.items()},
                "args": [show(v) for v in sig.args], "kwargs": {k: show(v) for k, v in sig.kwargs.items()}}

    calls = {
        "f0()": lambda: bind(f0), "f0(1)": lambda: bind(f0, 1), "f0(x=1)": lambda: bind(f0, x=1),
        "f1()": lambda: bind(f1), "f1(1)": lambda: bind(f1, 1), "f1('s')": lambda: bind(f1, "s"), "f1(a=1)": lambda: bind(f1, a=1),
        "f1(b=1)": lambda: bind(f1, b=1), "f1(1, 2)": lambda: bind(f1, 1, 2), "f1(True)": lambda: bind(f1, True),
        "f2(1)": lambda: bind(f2, 1), "f2(1, 'y')": lambda: bind(f2, 1, "y"), "f2(1, b=2)": lambda: bind(f2, 1, b=2),
        "f2(b='y')": lambda: bind(f2, b="y"), "f2(b='y', a=3)": lambda: bind(f2, b="y", a=3),
        "f3(1, 'x')": lambda: bind(f3, 1, "x"), "f3(None, 2.5)": lambda: bind(f3, None, 2.5), "f3(1)": lambda: bind(f3, 1),
        "f3([1], {'k': 1})": lambda: bind(f3, [1], {"k": 1}), "f3([], ())": lambda: bind(f3, [], ()),
        "f4()": lambda: bind(f4), "f4(1, 'a', 2.0)": lambda: bind(f4, 1, "a", 2.0), "f4(k=1)": lambda: bind(f4, k=1),
        "f5(1, 2, 3)": lambda: bind(f5, 1, 2, 3), "f5('x')": lambda: bind(f5, "x"), "f5(1, 'x')": lambda: bind(f5, 1, "x"),
        "f6(1, z=3)": lambda: bind(f6, 1, z=3), "f6(z=3)": lambda: bind(f6, z=3), "f6(a=1, z='s')": lambda: bind(f6, a=1, z="s"),
        "f7([1, 2])": lambda: bind(f7, [1, 2]), "f7([])": lambda: bind(f7, []), "f7(['a'])": lambda: bind(f7, ["a"]),
        "f7((1,))": lambda: bind(f7, (1,)),
        "f8(None)": lambda: bind(f8, None), "f8(3)": lambda: bind(f8, 3), "f8('s')": lambda: bind(f8, "s"),
        "f9(Derived())": lambda: bind(f9, Derived()), "f9(Base())": lambda: bind(f9, Base()), "f9(3)": lambda: bind(f9, 3),
        "f10('s', k=1)": lambda: bind(f10, "s", k=1), "f10(1.5)": lambda: bind(f10, 1.5), "f10(1, k='s')": lambda: bind(f10, 1, k="s"),
        "f11(1, k='k')": lambda: bind(f11, 1, k="k"), "f11(1, 2, 3, 4, k='k', z=0)": lambda: bind(f11, 1, 2, 3, 4, k="k", z=0),
        "f11(1)": lambda: bind(f11, 1),
        "f1(<int>)": lambda: bind(f1, px("e-int")), "f1(<str>)": lambda: bind(f1, px("e-str")),
        "f1(a=<int>)": lambda: bind(f1, a=px("e-int")), "f3(<int>, <str>)": lambda: bind(f3, px("e-int"), px("e-str")),
        "f7(<list>)": lambda: bind(f7, px("e-list")), "f9(<base>)": lambda: bind(f9, px("e-base")),
        "f4(<int>, 1)": lambda: bind(f4, px("e-int"), 1), "f10(<str>)": lambda: bind(f10, px("e-str")),
        "f1(<int, touched>)": lambda: bind(f1, px("e-int", touched=True)),
        "f3(<list, touched>, <base>)": lambda: bind(f3, px("e-list", touched=True), px("e-base")),
    }
    return {label: make() for label, make in calls.items()}
