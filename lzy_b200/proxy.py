"""Lazy proxies: an object that looks like the result of an @op but is computed at the last moment.

Same observable contract as the reference's ``lzy.proxy.proxy`` (pylzy/lzy/proxy/automagic.py:150-189
and pylzy/lzy/api/v1/utils/proxy_adapter.py:14-83):

* the proxy's *class* carries ``cls_attrs`` (``__lzy_proxied__``, ``__lzy_entry_id__`` …);
* touching any attribute or special method materialises the value exactly once; the value — or the
  exception raised while producing it — is cached on the proxy class;
* ``obj.__lzy_origin__`` forces materialisation and returns the real object,
  ``obj.__lzy_materialized__`` tells whether that already happened;
* special methods are looked up on the prototype types, so ``len(p)``, ``p == x``, ``for _ in p``,
  ``str(p)``, ``p + 1`` behave like the real value, and proxies passed as arguments to those methods
  are materialised first;
* ``type(p)`` is never the real type (``type(a) is str`` is False inside an op with
  ``lazy_arguments=True``, pylzy/tests/core/test_workflow.py:386-393) but ``p.__class__`` is, so
  ``isinstance(p, T)`` answers for the materialised value (pylzy/tests/proxy/test_proxy.py:48-71,280-284).

The implementation is a per-proxy class whose special methods are plain forwarding closures.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional, Sequence, Type

_SKIP = frozenset({
    "__class__", "__new__", "__init__", "__getattribute__", "__setattr__", "__delattr__", "__dict__", "__slots__",
    "__init_subclass__", "__subclasshook__", "__class_getitem__", "__weakref__", "__mro__", "__del__",
    "__reduce__", "__reduce_ex__", "__getstate__", "__setstate__", "__getnewargs__", "__getnewargs_ex__",
    "__sizeof__", "__doc__", "__module__", "__qualname__", "__annotations__", "__getattr__",
})


class _LazyMeta(type):
    """marker metaclass: ``is_proxy`` is an isinstance check on the class of the class"""


def is_proxy(obj: Any) -> bool:
    return isinstance(type(obj), _LazyMeta)


def _unwrap(x: Any) -> Any:
    return x.__lzy_origin__ if is_proxy(x) else x


def _resolve(cls) -> Any:
    state = type.__getattribute__(cls, "_lzy_state")
    if "value" in state:
        return state["value"]
    if "error" in state:
        raise state["error"]
    try:
        state["value"] = state["constructor"]()
    except Exception as e:  # cached: a second touch re-raises the same exception
        state["error"] = e
        raise
    return state["value"]


def _unwrap_tree(x: Any) -> Any:
    if is_proxy(x):
        return x.__lzy_origin__
    if isinstance(x, tuple):
        return tuple(_unwrap_tree(v) for v in x)
    if isinstance(x, list):
        return [_unwrap_tree(v) for v in x]
    if isinstance(x, dict):
        return {k: _unwrap_tree(v) for k, v in x.items()}
    return x


# Library dispatch protocols (`np.array_equal(p, a)`, `torch.equal(p, t)`, `np.add(p, 1)`): numpy and
# torch look these up on the *type* of the argument and pass the argument types along, so plain
# forwarding would hand the libraries a type they do not know.  Materialise and call the function.
def _array_function(self, func, types, args, kwargs):
    return func(*_unwrap_tree(args), **_unwrap_tree(kwargs))


def _array_ufunc(self, ufunc, method, *inputs, **kwargs):
    return getattr(ufunc, method)(*_unwrap_tree(inputs), **_unwrap_tree(kwargs))


def _torch_function(cls, func, types, args=(), kwargs=None):
    return func(*_unwrap_tree(args), **_unwrap_tree(kwargs or {}))


_PROTOCOLS = {"__array_function__": _array_function, "__array_ufunc__": _array_ufunc,
              "__torch_function__": classmethod(_torch_function)}


def _forwarder(name: str):
    def method(self, *args, **kwargs):
        target = _resolve(type(self))
        args = tuple(_unwrap(a) for a in args)
        kwargs = {k: _unwrap(v) for k, v in kwargs.items()}
        if name == "__bool__" and not hasattr(target, "__bool__"):
            return bool(target)
        return getattr(target, name)(*args, **kwargs)

    method.__name__ = name
    return method


_FORWARDERS: Dict[str, Any] = {}
_NS_CACHE: Dict[tuple, Dict[str, Any]] = {}


def _namespace_for(proto_types: tuple) -> Dict[str, Any]:
    """special-method forwarders for a tuple of prototype types (built once per tuple: creating a proxy
    is on the @op call path)"""
    ns = _NS_CACHE.get(proto_types)
    if ns is None:
        names = {"__bool__", "__eq__", "__ne__", "__hash__", "__str__", "__repr__"}
        for typ in proto_types:
            for name in dir(typ):
                if name.startswith("__") and name.endswith("__") and name not in _SKIP:
                    names.add(name)
        ns = {}
        for name in names:
            fwd = _PROTOCOLS.get(name)
            if fwd is None:
                fwd = _FORWARDERS.get(name)
                if fwd is None:
                    fwd = _FORWARDERS[name] = _forwarder(name)
            ns[name] = fwd
        _NS_CACHE[proto_types] = ns
    return ns


def proxy(constructor: Callable[[], Any], proto_types: Sequence[Type], cls_attrs: Optional[Dict[str, Any]] = None,
          obj_attrs: Optional[Dict[str, Any]] = None) -> Any:
    cls_attrs = dict(cls_attrs or {})
    obj_attrs = dict(obj_attrs or {})
    own = set(cls_attrs) | set(obj_attrs)
    ns: Dict[str, Any] = dict(_namespace_for(tuple(proto_types)))  # forwarders are shared, the dict is per class

    def __getattribute__(self, item):
        cls = type(self)
        if item == "__lzy_origin__":
            return _resolve(cls)
        if item == "__lzy_materialized__":
            return "value" in type.__getattribute__(cls, "_lzy_state")
        if item == "__class__":  # isinstance(p, T) asks this: answered by the real object (automagic.py:66-81)
            return type(_resolve(cls))
        if item in own or item in _PROTOCOLS:
            return object.__getattribute__(self, item)
        return getattr(_resolve(cls), item)

    def __setattr__(self, item, value):
        if item in own:
            return object.__setattr__(self, item, value)
        return setattr(_resolve(type(self)), item, value)

    def __del__(self):
        # A class is cyclic garbage by construction (class ↔ __dict__ ↔ mro), so the state hanging off it — the value, the
        # constructor's closure and whatever those pin: an HBM frame behind a tensor view, a pinned host block — would
        # wait for the cycle collector.  The proxy instance is not part of a cycle and dies with its last reference:
        # that is the moment to let go.  (Who asked for the value holds a reference of their own.)
        type.__getattribute__(type(self), "_lzy_state").clear()

    ns["__getattribute__"] = __getattribute__
    ns["__setattr__"] = __setattr__
    ns["__del__"] = __del__
    ns.update(cls_attrs)
    ns["_lzy_state"] = {"constructor": constructor}
    cls = _LazyMeta("LazyProxy", (object,), ns)
    inst = object.__new__(cls)
    for k, v in obj_attrs.items():
        object.__setattr__(inst, k, v)
    return inst
