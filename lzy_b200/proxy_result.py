"""The two outcomes of asking a snapshot for data: ``Result(value)`` — it is there — or ``Absence(cause)`` — it is
not (yet), with the reason if one is known.  The reference calls the pair ``Either`` and uses ``unwrap`` to turn an
absence into an ``AttributeError`` (pylzy/lzy/proxy/result.py; consumers: api/v1/snapshot.py:108-139,
api/v1/utils/proxy_adapter.py:55-83)."""
from __future__ import annotations

from typing import Any, Callable, Generic, Optional, TypeVar, Union

T = TypeVar("T")
U = TypeVar("U")


class Result(Generic[T]):
    """data that exists"""
    __slots__ = ("value",)
    __match_args__ = ("value",)

    def __init__(self, value: T):
        self.value = value

    def __bool__(self) -> bool:
        return True

    def __eq__(self, other: object) -> bool:
        return isinstance(other, Result) and self.value == other.value

    __hash__ = None  # type: ignore[assignment]

    def __repr__(self) -> str:
        return f"Result(value={self.value!r})"

    def map(self, fn: Callable[[T], U]) -> "Result[U]":
        return Result(fn(self.value))


class Absence:
    """data that does not exist; `cause` says why when somebody knows"""
    __slots__ = ("cause",)
    __match_args__ = ("cause",)

    def __init__(self, cause: Optional[Exception] = None):
        self.cause = cause

    def __bool__(self) -> bool:
        return False

    def __eq__(self, other: object) -> bool:
        return (isinstance(other, Absence) and type(self.cause) is type(other.cause)
                and getattr(self.cause, "args", None) == getattr(other.cause, "args", None))

    __hash__ = None  # type: ignore[assignment]

    def __repr__(self) -> str:
        return f"Absence(cause={self.cause!r})"

    def map(self, fn: Callable[[Any], Any]) -> "Absence":
        return self


Either = Union[Result, Absence]


def unwrap(outcome: Either) -> Any:
    """the value, or AttributeError naming the cause of the absence"""
    if outcome:
        return outcome.value  # type: ignore[union-attr]
    raise AttributeError(f"Cannot unwrap result, it is absent because of cause={outcome.cause!r}")  # type: ignore[union-attr]
