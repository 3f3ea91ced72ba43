"""``@op`` (the decorator of pylzy/lzy/core/op.py:18-61, same keyword surface): turns a function into a
``LazyCallWrapper``.  The result types come from ``output_types=`` or, failing that, from the return annotation — which
is therefore mandatory: an unannotated function is refused when it is decorated, not when it is first called."""
from __future__ import annotations

import inspect
from typing import Any, Callable, Optional, Sequence, TypeVar

from ..api.v1.utils.types import infer_return_type
from ..env import LzyEnvironment
from ..proxy_result import Absence
from .call import LazyCallWrapper

FuncT = TypeVar("FuncT", bound=Callable[..., Any])


def _where(f: Callable) -> str:
    """``function `name` at file:line`` — the wording of pylzy/lzy/utils/format.py:6-27, which the reference's tests
    match on (tests/core/test_calls.py:352-356)"""
    body = f.__func__ if inspect.ismethod(f) else f
    place = ""
    if inspect.isfunction(body):
        place = f" at {inspect.getfile(body)}"
        if hasattr(body, "__code__"):
            place += f":{body.__code__.co_firstlineno}"
    return f"{type(f).__name__} `{getattr(f, '__name__', repr(f))}`{place}"


def _declared_outputs(f: Callable) -> Sequence[type]:
    annotated = infer_return_type(f)
    if isinstance(annotated, Absence):
        raise TypeError(f"Return type is not annotated for {_where(f)}. Please for proper use of {op.__name__} "
                        f"annotate return type of your function.")
    return annotated.value


def op(func: Optional[FuncT] = None, *, env: Optional[LzyEnvironment] = None, output_types: Optional[Sequence[type]] = None,
       description: str = "", version: str = "0.0", cache: bool = False, lazy_arguments: bool = False):
    """usable bare (``@op``) and with options (``@op(cache=True, version="1.0")``)"""
    def wrap(f: Callable) -> LazyCallWrapper:
        outputs = output_types if output_types is not None else _declared_outputs(f)
        return LazyCallWrapper(f, outputs, env if env is not None else LzyEnvironment(), description, version, cache,
                               lazy_arguments)

    return wrap if func is None else wrap(func)
