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
