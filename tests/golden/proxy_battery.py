"""One battery of lazy-proxy probes, run twice: by tests/golden/make_golden.py against the *reference's own*
``lzy.proxy.proxy`` (pylzy/lzy/proxy/automagic.py, importable stand-alone in the authoring container) to produce
tests/golden/reference_proxy_vectors.json, and by tests/test_lazy_proxy.py against ``lzy_b200.proxy.proxy``, which
must give the same answers.  Every probe returns something JSON-able; an exception is recorded by type name."""
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Tuple


class Account:
    def __init__(self, owner, balance):
        self._owner, self.balance = owner, balance

    @property
    def owner(self):
        return self._owner

    def deposit(self, n):
        self.balance += n
        return self.balance

    @classmethod
    def kind(cls):
        return cls.__name__

    @staticmethod
    def join(a, b):
        return a + b

    def __len__(self):
        return self.balance

    def __contains__(self, item):
        return item == self._owner


@dataclass
class Point:
    x: int
    y: int = 0


def _probes(proxy: Callable[..., Any]) -> List[Tuple[str, Callable[[], Any]]]:
    def lazy(cls, *a, **k):
        return proxy(lambda: cls(*a, **k), (cls,))

    def none_of(cls):
        return proxy(lambda: None, (cls, type(None)))

    counter = {"n": 0}

    def counted():
        counter["n"] += 1
        return [1, 2, 3]

    def count_materialisations():
        p = proxy(counted, (list,))
        len(p), p[0], list(p), p == [1, 2, 3]
        return counter["n"]

    return [
        ("int_add", lambda: lazy(int, 40) + 2),
        ("int_radd", lambda: 2 + lazy(int, 40)),
        ("int_mul_proxies", lambda: lazy(int, 6) * lazy(int, 7)),
        ("int_floordiv_mod", lambda: [lazy(int, 17) // 5, lazy(int, 17) % 5]),
        ("int_neg_abs", lambda: [-lazy(int, 5), abs(lazy(int, -5))]),
        ("int_compare", lambda: [lazy(int, 3) < 4, lazy(int, 3) >= 4, lazy(int, 3) == 3, lazy(int, 3) != 3]),
        ("int_index_range", lambda: list(range(lazy(int, 4)))),
        ("int_hash_equals_origin", lambda: hash(lazy(int, 12345)) == hash(12345)),
        ("int_format", lambda: [f"{lazy(int, 255):x}", str(lazy(int, 7)), repr(lazy(int, 7))]),
        ("float_round", lambda: round(lazy(float, 2.567), 1)),
        ("bool_true_false", lambda: [bool(lazy(bool, True)), bool(lazy(bool, False)), bool(none_of(bool))]),
        ("str_methods", lambda: [lazy(str, "ab ").rstrip(), lazy(str, "ab").upper(), lazy(str, "a,b").split(",")]),
        ("str_concat_len_in", lambda: [lazy(str, "ab") + "c", len(lazy(str, "abc")), "b" in lazy(str, "abc")]),
        ("str_iter_getitem", lambda: [list(lazy(str, "xyz")), lazy(str, "xyz")[1], lazy(str, "xyz")[::-1]]),
        ("list_ops", lambda: [len(lazy(list, [1, 2])), lazy(list, [1, 2]) + [3], sorted(lazy(list, [3, 1, 2]))]),
        ("list_mutation", lambda: (lambda p: (p.append(4), p[-1], len(p)))(lazy(list, [1, 2, 3]))[1:]),
        ("dict_ops", lambda: (lambda p: [p["a"], sorted(p.keys()), "a" in p, p.get("z", 0)])(lazy(dict, {"a": 1, "b": 2}))),
        ("tuple_unpack", lambda: (lambda p: [a for a in p])(lazy(tuple, (1, "x")))),
        ("isinstance_declared", lambda: [isinstance(lazy(int, 1), int), isinstance(lazy(Account, "a", 1), Account)]),
        ("isinstance_none", lambda: [isinstance(none_of(int), int), isinstance(none_of(int), type(None))]),
        ("type_is_not_real_type", lambda: type(lazy(int, 1)) is int),
        ("dunder_class_is_real", lambda: lazy(int, 1).__class__ is int),
        ("origin_type", lambda: type(lazy(int, 1).__lzy_origin__).__name__),
        ("attr_property_method", lambda: (lambda p: [p.owner, p.balance, p.deposit(5), p.balance])(lazy(Account, "ann", 10))),
        ("attr_set_new_and_existing", lambda: (lambda p: (setattr(p, "tag", "t"), setattr(p, "balance", 1), [p.tag, p.balance]))(lazy(Account, "a", 0))[2]),
        ("classmethod_staticmethod", lambda: [lazy(Account, "a", 0).kind(), lazy(Account, "a", 0).join("x", "y")]),
        ("custom_len_contains", lambda: [len(lazy(Account, "ann", 3)), "ann" in lazy(Account, "ann", 3)]),
        ("dataclass_eq_repr", lambda: [lazy(Point, 1, 2) == Point(1, 2), repr(lazy(Point, 1, 2))]),
        ("none_attr", lambda: none_of(Account).owner),
        ("none_call_method", lambda: none_of(Account).deposit(1)),
        ("none_setattr", lambda: setattr(none_of(Account), "x", 1)),
        ("none_index", lambda: none_of(list)[0]),
        ("none_range", lambda: list(range(none_of(int)))),
        ("missing_attr", lambda: lazy(Account, "a", 0).nope),
        ("origin_exception_passes", lambda: lazy(dict, {})["missing"]),
        ("constructor_exception_passes", lambda: proxy(lambda: 1 // 0, (int,)) + 1),
        ("union_types_str", lambda: proxy(lambda: "s", (int, str)) + "t"),
        ("union_types_int", lambda: proxy(lambda: 10, (int, str)) + 1),
        ("materialised_once", count_materialisations),
        ("cls_attrs_visible", lambda: (lambda p: [p.__marker__, type(p).__marker__])(proxy(lambda: 1, (int,), cls_attrs={"__marker__": "m"}))),
    ]


def run(proxy: Callable[..., Any]) -> Dict[str, Any]:
    out: Dict[str, Any] = {}
    for name, probe in _probes(proxy):
        try:
            out[name] = probe()
        except Exception as e:  # noqa: BLE001 — the type of the failure is the recorded behaviour
            out[name] = "EXC:" + type(e).__name__
    return out
