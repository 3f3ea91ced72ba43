"""`lzy_b200.proxy.proxy` against the behaviours the reference pins for its lazy proxies
(pylzy/tests/proxy/test_proxy.py:48-305; implementation pylzy/lzy/proxy/automagic.py:66-173): isinstance by
the declared types, None results, attribute / descriptor / slot access through the origin, operators on
builtins, one materialisation, `__lzy_origin__`, C functions that insist on exact types."""
import json
from dataclasses import dataclass

import pytest

from lzy_b200.proxy import proxy


def lazy(cls, *args, **kwargs):
    return proxy(lambda: cls(*args, **kwargs), (cls,))


def lazy_none(cls):
    return proxy(lambda: None, (cls, type(None)))


class Account:
    def __init__(self, owner, balance, tag):
        self.__owner = owner
        self._balance = balance
        self.tag = tag

    @property
    def owner(self):
        return self.__owner

    @owner.setter
    def owner(self, value):
        self.__owner = value

    def get_tag(self):
        return self.tag

    @classmethod
    def kind(cls):
        return cls

    @staticmethod
    def join(left, right):
        return left + right


class Slotted:
    __slots__ = ("_members",)

    def __init__(self):
        self._members = set()


def test_isinstance_follows_the_declared_types_and_none_results():
    class A:
        pass

    assert isinstance(lazy(A), A) and isinstance(lazy(int), int)
    for cls in (A, int):
        p = lazy_none(cls)
        assert not isinstance(p, cls) and isinstance(p, type(None))


def test_attributes_properties_static_and_class_methods_go_to_the_origin():
    a = lazy(Account, "ann", 10, "t")
    assert (a.owner, a._balance, a.get_tag()) == ("ann", 10, "t")
    a.extra = 550
    assert a.extra == 550
    a.extra = 5
    assert a.extra == 5
    a.owner = "bob"
    assert a.owner == "bob"
    assert a.join("oh ", "hohoho") == "oh hohoho" and a.kind() is Account


def test_none_origin_has_none_of_the_declared_attributes():
    a = lazy_none(Account)
    for touch in (lambda: a._balance, lambda: a.owner, lambda: a.get_tag(), lambda: a.join("x", "y"), lambda: a.kind()):
        with pytest.raises(AttributeError):
            touch()
    with pytest.raises(AttributeError):
        a.extra = 550


def test_builtin_operators_and_protocols():
    n, acc = lazy(int, 42), lazy(int, 0)
    for _ in range(n):
        acc += 1
    assert acc == 42
    assert lazy(int, 10) + lazy(int, 50) == 60 and lazy(int, 10) * lazy(int, 50) == 500
    assert lazy(str, "string   ").rstrip() == "string" and len(lazy(list, [1, 2, 3, 4, 5])) == 5
    with pytest.raises(AttributeError):
        for _ in range(lazy_none(int)):
            pass


def test_truthiness():
    assert lazy(bool, True) and not lazy_none(bool)
    assert proxy(lambda: "str", (str, type(None))) and not lazy_none(str)


def test_slots_are_reachable_and_assignable():
    s = lazy(Slotted)
    assert isinstance(s._members, set)
    s._members = None
    assert s._members is None and s.__slots__ == Slotted.__slots__
    with pytest.raises(AttributeError):
        lazy_none(Slotted)._members


class _Field:
    def __init__(self, value):
        self.value = value

    def __get__(self, instance, owner):
        return self.value

    def __eq__(self, other):
        return other.value == self.value


class _SettableField(_Field):
    def __set__(self, instance, value):
        self.value = value


def test_descriptors_read_and_write_with_a_single_materialisation():
    class Model:
        a = _Field(41)
        b = _Field(42)

    hits = []
    p = proxy(lambda: hits.append(1) or Model(), (Model,))
    assert p.a != p.b and len(hits) == 1
    with pytest.raises(AttributeError):
        lazy_none(Model).a

    class Mutable:
        a = _SettableField(41)
        b = _SettableField(42)

    hits.clear()
    p = proxy(lambda: hits.append(1) or Mutable(), (Mutable,))
    p.a = 42
    assert len(hits) == 1 and p.a == p.b
    with pytest.raises(AttributeError):
        lazy_none(Mutable).a = 42


def test_exceptions_of_the_origin_pass_through():
    class Boom(Exception):
        pass

    class Thing:
        def go(self):
            raise Boom("Ahaa")

    with pytest.raises(Boom):
        proxy(Thing, (Thing,)).go()


def test_items_of_containers():
    @dataclass
    class A:
        val: int

    p = proxy(lambda: [A(0), A(1), A(2)], (list,))
    assert [p[i].val for i in range(3)] == [0, 1, 2]
    with pytest.raises(AttributeError):
        lazy_none(list)[0].val


def test_class_with_custom_new_and_optional_results():
    class Custom:
        def __new__(cls):
            return super().__new__(cls)

        def __init__(self):
            self.attr = 42

    assert proxy(Custom, (Custom,)).attr == 42

    @dataclass
    class A:
        a: int

    assert proxy(lambda: A(0), (A, type(None))).a == 0
    with pytest.raises(AttributeError):
        lazy_none(str).a


def test_origin_and_exact_type_consumers():
    n = lazy(int, 42)
    assert n.__class__ is int and type(n) is not int and type(n.__lzy_origin__) is int
    s = lazy(str, "[]")
    with pytest.raises(TypeError):
        json.dumps(object.__new__(type(s))) if False else json.loads(s)  # C scanner wants a real str …
    assert json.loads(s.__lzy_origin__) == []  # … which the origin is


def test_union_of_declared_types():
    p1, p2 = proxy(lambda: "str", (int, str)), proxy(lambda: 10, (int, str))
    assert p1 + "str" == "strstr"
    assert isinstance(p2, int) and sum(1 for _ in range(p2)) == 10


def test_same_answers_as_the_reference_proxy_on_the_shared_battery():
    """tests/golden/reference_proxy_vectors.json was produced by running tests/golden/proxy_battery.py against the
    reference's own `lzy.proxy.proxy` (pylzy/lzy/proxy/automagic.py) in the authoring container"""
    import json
    from pathlib import Path

    from tests.golden import proxy_battery
    golden = json.loads((Path(__file__).parent / "golden" / "reference_proxy_vectors.json").read_text())
    ours = json.loads(json.dumps(proxy_battery.run(proxy), default=str))
    assert set(ours) == set(golden)
    assert {k: ours[k] for k in golden if ours[k] != golden[k]} == {}


def test_entry_proxies_obtain_their_value_as_the_reference_adapter_does():
    """tests/golden/reference_materialize_vectors.json: pylzy/lzy/api/v1/utils/proxy_adapter.py's own lzy_proxy (executed
    in the authoring container) over a fake workflow whose snapshot answers are scripted — a supplied value asks
    nobody; a stored one is read without a barrier; otherwise exactly one barrier and a second read; an Absence with
    a cause raises the cause, one without raises RuntimeError naming entry and workflow; the markers (is proxy,
    entry id, materialized) compute nothing; a second touch reads nothing again"""
    import importlib.util
    import json
    from pathlib import Path

    from lzy_b200.api.v1.utils import proxy_adapter as adapter
    from lzy_b200.proxy_result import Absence, Result

    golden = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("materialize_battery", golden / "materialize_battery.py")
    battery = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(battery)
    ours = battery.run(dict(lzy_proxy=adapter.lzy_proxy, materialize=adapter.materialize, materialized=adapter.materialized,
                            is_lzy_proxy=adapter.is_lzy_proxy, get_proxy_entry_id=adapter.get_proxy_entry_id,
                            Result=Result, Absence=Absence))
    assert json.loads(json.dumps(ours)) == json.loads((golden / "reference_materialize_vectors.json").read_text())


def test_a_dead_proxy_lets_go_of_its_value_without_the_cycle_collector():
    """the proxy's class is cyclic garbage by construction; what it caches (a tensor view pinning an HBM frame, a pinned
    host block) must not wait for the collector to find it"""
    import gc
    import weakref

    class Payload:
        pass

    gc.disable()
    try:
        made = []

        def build():
            made.append(Payload())
            return made[-1]

        p = proxy(build, (Payload,))
        value = p.__lzy_origin__
        ref = weakref.ref(value)
        made.clear()
        del value
        assert ref() is not None       # cached on the proxy while the proxy lives
        del p
        assert ref() is None           # gone with it — no gc.collect() in between
    finally:
        gc.enable()
