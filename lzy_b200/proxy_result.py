"""Either / Result / Absence (pylzy/lzy/proxy/result.py:1-24)."""
from dataclasses import dataclass
from typing import Generic, TypeVar, Union

T = TypeVar("T")


@dataclass
class Absence:
    cause: Union[Exception, None] = None


@dataclass
class Result(Generic[T]):
    value: T


Either = Union[Result, Absence]


def unwrap(res):
    if isinstance(res, Absence):
        raise AttributeError(f"Cannot unwrap result, it is absent because of cause={res.cause!r}")
    return res.value
