"""Which ops of one submitted graph actually run, and who supplies what.

The reference keys a graph by *data URI*, not by call: every output slot URI has exactly one supplier op, an op whose
outputs are all supplied by earlier ops of the same graph is dropped, an op that would supply only part of what is
already supplied is an error (lzy/lzy-service/src/main/java/ai/lzy/service/dao/DataFlowGraph.java:51-72), and
consumers are wired to suppliers through the URI (:74-105).  Two identical ``cache=True`` calls in one barrier therefore
run the op once; both results read the same object.  Before that, an op whose results are all in storage already is
not submitted at all.  Both runtimes of this repository (the HBM runtime and the CPU twin used by the tests) plan their
graphs with this function, so the rule is checked without a GPU."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import TYPE_CHECKING, Callable, Dict, List, Sequence

if TYPE_CHECKING:
    from .call import LzyCall


@dataclass
class DataflowPlan:
    calls: List["LzyCall"] = field(default_factory=list)       # the ops that run, in submission order
    cached: List["LzyCall"] = field(default_factory=list)      # dropped: every result is in storage already
    duplicates: List["LzyCall"] = field(default_factory=list)  # dropped: an earlier op of this graph supplies the same URIs
    supplier: Dict[str, int] = field(default_factory=dict)     # result entry id (of kept AND duplicate calls) -> index in `calls`
    also_supplies: Dict[int, List[str]] = field(default_factory=dict)  # index in `calls` -> entry ids of duplicates it stands for

    @property
    def dropped(self) -> List["LzyCall"]:
        return self.cached + self.duplicates

    def outputs_of(self, index: int) -> List[str]:
        """every entry that exists once op `index` has finished"""
        return list(self.calls[index].entry_ids) + self.also_supplies.get(index, [])


def plan_dataflow(calls: Sequence["LzyCall"], uri_of: Callable[[str], str], exists: Callable[[str], bool]) -> DataflowPlan:
    plan = DataflowPlan()
    supplied: Dict[str, int] = {}  # storage uri -> index in plan.calls
    for call in calls:
        uris = [uri_of(eid) for eid in call.entry_ids]
        if call.cache and uris and all(exists(u) for u in uris):
            plan.cached.append(call)
            continue
        if uris and all(u in supplied for u in uris):
            plan.duplicates.append(call)
            for eid, u in zip(call.entry_ids, uris):
                plan.supplier[eid] = supplied[u]
                plan.also_supplies.setdefault(supplied[u], []).append(eid)
            continue
        for u in uris:
            if u in supplied:
                raise RuntimeError(f"Output slot with uri '{u}' already exists")
        index = len(plan.calls)
        plan.calls.append(call)
        for eid, u in zip(call.entry_ids, uris):
            supplied[u] = index
            plan.supplier[eid] = index
    return plan


def sole_consumer_outputs(plan: DataflowPlan, inputs_of: Sequence[Sequence[str]], placement: Sequence[int],
                          claims_of: Callable[[str], int], excluded: Callable[[str], bool]) -> set:
    """Result entries of this graph whose frame can be HANDED to their one consumer instead of being copied for it.

    `FileInputBackend.toOutput()` moves a file by ownership (lzy/slots/.../backends/FileInputBackend.java:38-45); an
    HBM frame may travel the same way along a producer → consumer edge exactly when nobody else can ever ask for it:

    * one op of this graph consumes the entry, on the producer's GPU (`placement`), and no dropped duplicate shares it;
    * `excluded(eid)` is false — the runtime's veto for cached (`cache=True`) results and fan-out deliveries;
    * `claims_of(eid)` — the snapshot's pin count: one per queued consumer call and one while the result's lazy proxy
      is alive — leaves nobody but that consumer: either 1 (the proxy is gone), or 2 with the proxy kept alive by
      nothing but the consumer call's own argument list (``y = f(y)`` in a loop).  A proxy the caller still holds —
      in a variable, a list, another object — can be materialized after the graph, so its frame stays in the store and
      the consumer gets a private copy (snapshot-at-call-time, pylzy/tests/api/v1/test_entries.py:331-384)."""
    import sys

    from ..api.v1.utils.proxy_adapter import get_proxy_entry_id, is_lzy_proxy
    consumers: Dict[str, List[int]] = {}
    for i, ins in enumerate(inputs_of):
        for eid in ins:
            consumers.setdefault(eid, []).append(i)
    shared = {eid for eids in plan.also_supplies.values() for eid in eids}
    for index, eids in plan.also_supplies.items():
        if eids:
            shared.update(plan.calls[index].entry_ids)
    out = set()
    for index, call in enumerate(plan.calls):
        for eid in call.entry_ids:
            cons = consumers.get(eid, ())
            if len(cons) != 1 or eid in shared or excluded(eid) or placement[cons[0]] != placement[index]:
                continue
            claims = claims_of(eid)
            if claims == 1:
                out.add(eid)
                continue
            if claims != 2:
                continue
            consumer = plan.calls[cons[0]]
            held = [a for a in (*consumer.args, *consumer.kwargs.values()) if is_lzy_proxy(a) and get_proxy_entry_id(a) == eid]
            # references to the proxy right now: the consumer's argument tuple (or kwargs dict), `held`, and getrefcount's
            # own argument.  One more than that is somebody who can still materialize it.
            if len(held) == 1 and sys.getrefcount(held[0]) == 3:
                out.add(eid)
            del held
    return out
