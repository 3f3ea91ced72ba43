"""Runtime type inference used for @op argument / result validation.

The reference takes ``get_type`` and ``EmptyContent`` from the un-vendored ``serialzy.types``
(call sites: pylzy/lzy/core/call.py:12-14,307,326; pylzy/lzy/api/v1/startup.py:15,63); the behaviour
restated here is the one its tests rely on (pylzy/tests/core/test_workflow.py:73-128,290-330 and
pylzy/lzy/api/v1/utils/types.py:76-79): containers are typed by their first element, empty containers
are typed with the ``EmptyContent`` marker, which is compatible with any element type."""
from __future__ import annotations

from typing import Any, Dict, List, Tuple, Type


class EmptyContent:
    """element type of an empty list / tuple / dict"""


LARGE_TUPLE = 64


def get_type(obj: Any) -> Type:
    from ..proxy import is_proxy
    if is_proxy(obj):
        # the result of another op placed inside a local container (`f([a, b, c])`, tests/core/test_workflow.py:435-447)
        # is described by the value it stands for; asking materialises it, as touching it in any other way would
        return get_type(obj.__lzy_origin__)
    if type(obj) is list:
        if len(obj) == 0:
            return List[EmptyContent]  # type: ignore[return-value]
        return List[get_type(obj[0])]  # type: ignore[return-value,misc]
    if type(obj) is tuple:
        if len(obj) == 0:
            return Tuple[EmptyContent]  # type: ignore[return-value]
        if len(obj) > LARGE_TUPLE:
            # a long tuple is described by its first element, like a list (tests/core/test_calls.py:142-149 pins
            # 3 elements → Tuple[int, int, int] and 1000 → Tuple[int, ...]; the cut-over lives in serialzy, which
            # is not in the reference tree — 64 here)
            return Tuple[get_type(obj[0]), ...]  # type: ignore[return-value,misc]
        return Tuple[tuple(get_type(o) for o in obj)]  # type: ignore[return-value,misc]
    if type(obj) is dict:
        if len(obj) == 0:
            return Dict[EmptyContent, EmptyContent]  # type: ignore[return-value]
        k, v = next(iter(obj.items()))
        return Dict[get_type(k), get_type(v)]  # type: ignore[return-value,misc]
    return type(obj)
