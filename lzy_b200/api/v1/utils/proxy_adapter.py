"""Lazy proxies that stand for snapshot entries — the functions of pylzy/lzy/api/v1/utils/proxy_adapter.py:14-83
(``lzy_proxy``, ``is_lzy_proxy``, ``get_proxy_entry_id``, ``materialize``, ``materialized``) over this package's
``proxy``.

The proxy *class* is tagged (every proxy has a class of its own, so class attributes are per proxy and asking them
never touches — i.e. never materialises — the origin).  Touching the proxy runs `EntryOrigin`: the value the caller
supplied, else what the snapshot has, else one barrier of the owning workflow and what the snapshot has then."""
from __future__ import annotations

from typing import TYPE_CHECKING, Any, Sequence, Type

from ....proxy import proxy
from ....proxy_result import Absence, Either, Result
from ....utils.event_loop import LzyEventLoop

if TYPE_CHECKING:
    from ....core.workflow import LzyWorkflow

_TAG = "__lzy_proxied__"
_ENTRY = "__lzy_entry_id__"


def is_lzy_proxy(obj: Any) -> bool:
    return getattr(type(obj), _TAG, False) is True


def get_proxy_entry_id(obj: Any) -> str:
    if is_lzy_proxy(obj):
        return str(getattr(type(obj), _ENTRY))
    raise ValueError(f"Object {obj} is not a lazy proxy")


def materialized(obj: Any) -> bool:
    """has the origin been computed already?  (asking does not compute it)"""
    return bool(obj.__lzy_materialized__)


def materialize(obj: Any) -> Any:
    """the real object behind the proxy, computing it if need be"""
    return obj.__lzy_origin__


class EntryOrigin:
    """how one entry's value is obtained when its proxy is first touched"""

    def __init__(self, entry_id: str, workflow: "LzyWorkflow", known: Either):
        self.entry_id, self.workflow, self.known = entry_id, workflow, known

    async def fetch(self) -> Any:
        if isinstance(self.known, Result):
            return self.known.value
        snapshot = self.workflow.snapshot
        stored = await snapshot.get_data(self.entry_id)
        if not isinstance(stored, Result):
            await self.workflow._barrier()  # the entry is the result of a queued call: run the queue once
            stored = await snapshot.get_data(self.entry_id)
        if isinstance(stored, Result):
            return stored.value
        if stored.cause:
            raise stored.cause
        raise RuntimeError(f"Cannot materialize data with entry id {self.entry_id} from workflow {self.workflow.name}")

    def __call__(self) -> Any:
        inline = getattr(self.workflow, "_inline", None)  # (test doubles stand in for workflows: no such method, no fast path)
        return LzyEventLoop.run_async(self.fetch(), inline=bool(inline()) if callable(inline) else False)


def lzy_proxy(entry_id: str, types: Sequence[Type], wflow: "LzyWorkflow", value: Either = Absence()) -> Any:
    return proxy(EntryOrigin(entry_id, wflow, value), types, cls_attrs={_TAG: True, _ENTRY: entry_id})
