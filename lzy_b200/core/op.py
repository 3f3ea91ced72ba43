"""The ``@op`` decorator (pylzy/lzy/core/op.py:18-61): infers the output types from the return
annotation (mandatory — ``TypeError`` at decoration time otherwise) and replaces the function with a
``LazyCallWrapper``."""
from __future__ import annotations

from typing import Any, Callable, Optional, Sequence, TypeVar

from ..api.v1.utils.types import infer_return_type
from ..env import LzyEnvironment
from ..proxy_result import Absence
from .call import LazyCallWrapper

FuncT = TypeVar("FuncT", bound=Callable[..., Any])


def _pretty(f: Callable) -> str:
    return f"{getattr(f, '__module__', '?')}.{getattr(f, '__qualname__', repr(f))}"


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
