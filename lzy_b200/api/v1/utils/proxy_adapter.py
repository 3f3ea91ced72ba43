"""Proxies bound to snapshot entries (pylzy/lzy/api/v1/utils/proxy_adapter.py:14-83)."""
from __future__ import annotations

from typing import TYPE_CHECKING, Any, Sequence, Type, cast

from ....proxy import proxy
from ....proxy_result import Absence, Either, Result
from ....utils.event_loop import LzyEventLoop

if TYPE_CHECKING:
    from ....core.workflow import LzyWorkflow

_PROXIED = "__lzy_proxied__"
_ENTRY_ID = "__lzy_entry_id__"


def is_lzy_proxy(obj: Any) -> bool:
    cls = type(obj)
    return bool(getattr(cls, _PROXIED, False))


def get_proxy_entry_id(obj: Any) -> str:
    if not is_lzy_proxy(obj):
        raise ValueError(f"Object {obj} is not a lazy proxy")
    return cast(str, getattr(type(obj), _ENTRY_ID))


def materialized(obj: Any) -> bool:
    return obj.__lzy_materialized__  # type: ignore[no-any-return]


def materialize(obj: Any) -> Any:
    return obj.__lzy_origin__


def lzy_proxy(entry_id: str, types: Sequence[Type], wflow: "LzyWorkflow", value: Either = Absence()) -> Any:
    async def __materialize() -> Any:
        if isinstance(value, Result):
            return value.value
        data = await wflow.snapshot.get_data(entry_id)
        if isinstance(data, Result):
            return data.value
        await wflow._barrier()
        new_data = await wflow.snapshot.get_data(entry_id)
        if isinstance(new_data, Result):
            return new_data.value
        if new_data.cause:
            raise new_data.cause
        raise RuntimeError(f"Cannot materialize data with entry id {entry_id} from workflow {wflow.name}")

    return proxy(lambda: LzyEventLoop.run_async(__materialize()), types, cls_attrs={_PROXIED: True, _ENTRY_ID: entry_id})
