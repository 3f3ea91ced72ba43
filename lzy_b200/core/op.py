"""The ``@op`` decorator (pylzy/lzy/core/op.py:18-61): infers the output types from the return
annotation (mandatory — ``TypeError`` at decoration time otherwise) and replaces the function with a
``LazyCallWrapper``."""
from __future__ import annotations

import inspect

from typing import Any, Callable, Optional, Sequence, TypeVar

from ..api.v1.utils.types import infer_return_type
from ..env import LzyEnvironment
from ..proxy_result import Absence
from .call import LazyCallWrapper

FuncT = TypeVar("FuncT", bound=Callable[..., Any])


def _pretty(f: Callable) -> str:
    """"function `name` at file:line" — the wording of pylzy/lzy/utils/format.py:6-27, which the reference's tests
    match on (tests/core/test_calls.py:352-356)"""
    text = f"{type(f).__name__} `{getattr(f, '__name__', repr(f))}`"
    target = f.__func__ if inspect.ismethod(f) else f
    if inspect.isfunction(target):
        text += f" at {inspect.getfile(target)}"
        if hasattr(target, "__code__"):
            text += f":{target.__code__.co_firstlineno}"
    return text


def op(func: Optional[FuncT] = None, *, env: Optional[LzyEnvironment] = None, output_types: Optional[Sequence[type]] = None,
       description: str = "", version: str = "0.0", cache: bool = False, lazy_arguments: bool = False):
    def deco(f):
        nonlocal output_types
        out = output_types
        if out is None:
            inferred = infer_return_type(f)
            if isinstance(inferred, Absence):
                raise TypeError(f"Return type is not annotated for {_pretty(f)}. Please for proper use of {op.__name__} "
                                f"annotate return type of your function.")
            out = inferred.value
        return LazyCallWrapper(function=f, output_types=out, env=env or LzyEnvironment(), description=description,
                               version=version, cache=cache, lazy_arguments=lazy_arguments)

    if func is None:
        return deco
    return deco(func)
