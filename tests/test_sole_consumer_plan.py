"""Which op outputs may be HANDED to their consumer instead of copied for it (lzy_b200/core/dataflow.py:
sole_consumer_outputs) — the planning half of VERDICT r1 #3, checked without a GPU on calls built by the real API front
over a pin-counting snapshot.  The rule under test is snapshot-at-call-time (pylzy/tests/api/v1/test_entries.py:331-384):
whoever can still ask for a result keeps it."""
from typing import Dict, List

import pytest

from lzy_b200.api.v1 import Lzy, op
from lzy_b200.core.dataflow import plan_dataflow, sole_consumer_outputs
from tests.mocks import MemSnapshot, RuntimeMock


class PinSnapshot(MemSnapshot):
    def __init__(self, *a):
        super().__init__(*a)
        self.pins: Dict[str, int] = {}
        self.persistent = set()

    def pin(self, eid):
        self.pins[eid] = self.pins.get(eid, 0) + 1

    def unpin(self, eid):
        self.pins[eid] = self.pins.get(eid, 0) - 1

    def persist(self, eid):
        self.persistent.add(eid)

    def claims(self, eid):
        return self.pins.get(eid, 0)


class PlanningRuntime(RuntimeMock):
    """records, for every submitted graph, which outputs the planner would hand over (nothing is executed)"""

    def __init__(self, placement=None, fanout=()):
        super().__init__(execute=False)
        self.placement_of = placement or (lambda i, call: 0)
        self.fanout = set(fanout)
        self.decisions: List[dict] = []

    def snapshot(self, workflow, prefix):
        return PinSnapshot(workflow.owner.serializer_registry, prefix, self.blobs, self.store_counts)

    async def exec(self, calls, progress):
        snap = self.workflow.snapshot
        for c in calls:  # results need URIs for the planner
            for eid in c.entry_ids:
                snap.get(eid).storage_uri = f"mem://ops/{c.callable_name}/{c.id}/{eid}"
        plan = plan_dataflow(calls, lambda eid: snap.get(eid).storage_uri, lambda uri: False)
        inputs_of = [list(dict.fromkeys(list(c.arg_entry_ids) + list(c.kwarg_entry_ids.values()))) for c in plan.calls]
        placement = [self.placement_of(i, c) for i, c in enumerate(plan.calls)]
        sole = sole_consumer_outputs(plan, inputs_of, placement, snap.claims,
                                     lambda eid: eid in self.fanout or eid in snap.persistent)
        self.decisions.append({"calls": plan.calls, "sole": sole})


@op
def inc(x: int) -> int:
    return x + 1


@op
def add(a: int, b: int) -> int:
    return a + b


@op(cache=True, version="1")
def inc_cached(x: int) -> int:
    return x + 1


def names(decision):
    """call index → whether its (only) output is handed over"""
    return [c.entry_ids[0] in decision["sole"] for c in decision["calls"]]


def test_a_rebinding_chain_hands_every_intermediate_over_and_keeps_the_last():
    rt = PlanningRuntime()
    with Lzy(runtime=rt).workflow("wf"):
        y = 1
        for _ in range(4):
            y = inc(y)  # the old proxy lives on only in the next call's argument list
    assert names(rt.decisions[0]) == [True, True, True, False]  # the caller still holds the last proxy


def test_a_proxy_the_caller_still_holds_keeps_its_frame():
    rt = PlanningRuntime()
    with Lzy(runtime=rt).workflow("wf"):
        a = inc(1)
        b = inc(a)        # `a` stays in scope: it can be materialized after the graph
        c = inc(b)
        del b             # nobody can ask for b any more
        d = inc(c)
        kept = [c]        # a container counts like a variable
        del c
    assert names(rt.decisions[0]) == [False, True, False, False]
    assert kept and a is not None and d is not None


def test_two_consumers_or_another_gpu_or_a_fanout_or_a_cached_result_are_copied():
    rt = PlanningRuntime()
    with Lzy(runtime=rt).workflow("wf"):
        s = inc(1)
        t = add(inc(s), inc(s))  # s has two consumers
        del s, t
    first = rt.decisions[0]
    by_name = [(c.callable_name, c.entry_ids[0] in first["sole"]) for c in first["calls"]]
    assert by_name == [("inc", False), ("inc", True), ("inc", True), ("add", False)]

    rt = PlanningRuntime(placement=lambda i, call: i % 2)  # neighbours on different GPUs: never a hand-over
    with Lzy(runtime=rt).workflow("wf"):
        y = inc(inc(inc(1)))
        del y
    assert names(rt.decisions[0]) == [False, False, False]

    rt = PlanningRuntime()
    with Lzy(runtime=rt).workflow("wf"):
        y = inc(inc_cached(inc(1)))
        del y
    assert names(rt.decisions[0]) == [True, False, False]  # the cached result stays in the store for later runs


def test_the_same_proxy_passed_twice_is_not_handed_over():
    rt = PlanningRuntime()
    with Lzy(runtime=rt).workflow("wf"):
        y = inc(1)
        z = add(y, y)
        del y, z
    assert names(rt.decisions[0]) == [False, False]
