"""One battery of questions about annotations, asked twice: by tests/golden/make_golden.py of the *reference's own*
pylzy/lzy/api/v1/utils/types.py (loaded by file path in the authoring container, its absent third party `serialzy`
replaced by two-name stubs) to produce tests/golden/reference_types_vectors.json, and by tests/test_call_signatures.py
of lzy_b200/api/v1/utils/types.py, which must answer the same.  `ns` carries the five functions under test plus the
module's own ``EmptyContent`` / ``Absence`` / ``Result``; every answer is JSON-able."""
from typing import Any, Dict, List, Optional, Tuple, Union


class Base:
    pass


class Derived(Base):
    pass


def _name(t) -> str:
    return getattr(t, "__name__", None) or repr(t)


def _names(ts) -> list:
    return [_name(t) for t in ts]


class _ByType:
    def __init__(self, typ):
        self._typ = typ

    def supported_types(self):
        return self._typ


class _ByPredicate:
    def __init__(self, accepted):
        self._accepted = accepted

    def supported_types(self):
        return lambda t: t in self._accepted


class FakeRegistry:
    """int → a by-type serializer; Base → a by-type serializer for Base; everything else → a predicate serializer"""

    def find_serializer_by_type(self, typ):
        if typ is int:
            return _ByType(int)
        if typ is Base:
            return _ByType(Base)
        return _ByPredicate({str, list, bool, Derived})


def f_plain(a, b=2, *, c="x", d):
    return a


def f_none():
    pass


def f_int() -> int:
    return 1


def f_literal_pair() -> (int, str):
    return 1, ""


def f_typing_pair() -> Tuple[int, str]:
    return 1, ""


def f_returns_none() -> None:
    return None


def f_string_annotation() -> "List[int]":
    return []


def f_optional() -> Optional[Base]:
    return None


def run(ns: Dict[str, Any]) -> Dict[str, Any]:
    EMPTY = ns["EmptyContent"]
    out: Dict[str, Any] = {}

    real = ns["infer_real_types"]
    out["infer_real_types"] = {label: _names(real(t)) for label, t in {
        "int": int, "List[int]": List[int], "Optional[str]": Optional[str],
        "Union[int, List[str], None]": Union[int, List[str], None], "Dict[str, int]": Dict[str, int],
        "Tuple[int, ...]": Tuple[int, ...], "NoneType": type(None), "Base": Base,
        "Union[Base, Dict[str, Derived]]": Union[Base, Dict[str, Derived]], "Optional[Union[int, str]]": Optional[Union[int, str]],
    }.items()}

    def returned(f):
        r = ns["infer_return_type"](f)
        return "Absence" if isinstance(r, ns["Absence"]) else _names(r.value)

    out["infer_return_type"] = {f.__name__: returned(f) for f in
                                (f_none, f_int, f_literal_pair, f_typing_pair, f_returns_none, f_string_annotation, f_optional)}
    out["get_default_args"] = {f.__name__: ns["get_default_args"](f) for f in (f_plain, f_int)}

    sub = ns["is_subtype"]
    pairs = {
        "int<=int": (int, int), "bool<=int": (bool, int), "int<=float": (int, float), "int<=str": (int, str),
        "List[int]<=List": (List[int], List), "List[int]<=List[int]": (List[int], List[int]),
        "List[int]<=List[str]": (List[int], List[str]), "List<=List[int]": (List, List[int]),
        "List[EMPTY]<=List[int]": (List[EMPTY], List[int]), "List[EMPTY]<=List": (List[EMPTY], List),
        "List[EMPTY]<=Dict[str, int]": (List[EMPTY], Dict[str, int]), "List[EMPTY]<=int": (List[EMPTY], int),
        "Tuple[EMPTY]<=Tuple[int, ...]": (Tuple[EMPTY], Tuple[int, ...]), "Tuple[EMPTY]<=Tuple[int, str]": (Tuple[EMPTY], Tuple[int, str]),
        "Tuple[EMPTY]<=List[int]": (Tuple[EMPTY], List[int]),
        "Dict[EMPTY, EMPTY]<=Dict[str, int]": (Dict[EMPTY, EMPTY], Dict[str, int]), "Dict[EMPTY, EMPTY]<=List[int]": (Dict[EMPTY, EMPTY], List[int]),
        "int<=Optional[int]": (int, Optional[int]), "NoneType<=Optional[int]": (type(None), Optional[int]),
        "str<=Union[int, str]": (str, Union[int, str]), "float<=Union[int, str]": (float, Union[int, str]),
        "int<=Any": (int, Any), "Derived<=Base": (Derived, Base), "Base<=Derived": (Base, Derived),
        "List[Derived]<=List[Base]": (List[Derived], List[Base]), "Dict[str, int]<=Dict[str, Any]": (Dict[str, int], Dict[str, Any]),
    }
    out["is_subtype"] = {label: bool(sub(a, b)) for label, (a, b) in pairs.items()}

    compatible, registry = ns["check_types_serialization_compatible"], FakeRegistry()
    cases = {
        "int/int": (int, int), "int/bool": (int, bool), "str/str": (str, str), "str/int": (str, int),
        "Union[int, str]/str": (Union[int, str], str), "Union[int, str]/float": (Union[int, str], float),
        "Optional[int]/int": (Optional[int], int), "List[int]/list": (List[int], list), "List[int]/tuple": (List[int], tuple),
        "Base/Base": (Base, Base), "Base/Derived": (Base, Derived), "Derived/Derived": (Derived, Derived),
        "Optional[Base]/Derived": (Optional[Base], Derived), "float/float": (float, float),
    }
    out["check_types_serialization_compatible"] = {label: bool(compatible(a, t, registry)) for label, (a, t) in cases.items()}
    return out
