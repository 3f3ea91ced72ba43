"""Graph dispatcher vs the reference's BFS processor.

Known-answer scenarios restate lzy/graph-executor/src/test/java/ai/lzy/graph/test/GraphExecutorTest.java
(testSimple :97-137, testErrorInTask :140-163, testStop :165-188, testCycles :190-218,
testRhombus :291-326, testOneVertex); the randomized test drives the C++ dispatcher and the
oracle's literal restatement of BfsGraphProcessor in lockstep and compares the executing sets."""
import random

import pytest

from lzy_b200 import _host as H
from lzy_b200._host import GraphDispatcher, GrpcStatusError
from oracle.graph_bfs import BfsOracle, T_ERROR, T_SUCCESS


def names(xs):
    return sorted(str(x + 1) for x in xs)


def build(vertices, edges):
    return GraphDispatcher(vertices, [(a - 1, b - 1) for a, b in edges])


SIMPLE_EDGES = [(1, 2), (3, 2), (3, 4), (5, 6), (7, 6), (7, 8), (9, 10), (10, 8)]


def done(g, *ids):
    for i in ids:
        g.task_done(int(i) - 1)


def test_simple_frontier_order():
    g = build(10, SIMPLE_EDGES)
    assert names(g.next()) == ["1", "3", "5", "7", "9"]
    assert names(g.executing()) == ["1", "3", "5", "7", "9"]
    done(g, 1, 5, 7, 9)
    g.next()
    assert names(g.executing()) == ["10", "3", "6"]
    done(g, 6, 10)
    g.next()
    assert names(g.executing()) == ["3", "8"]
    done(g, 3, 8)
    g.next()
    assert names(g.executing()) == ["2", "4"]
    done(g, 2, 4)
    assert g.next() == [] and g.status == H.G_COMPLETED


def test_rhombus():
    g = build(4, [(1, 2), (1, 3), (2, 4), (3, 4)])
    assert names(g.next()) == ["1"]
    done(g, 1)
    assert names(g.next()) == ["2", "3"]
    done(g, 3)
    assert g.next() == [] and names(g.executing()) == ["2"]
    done(g, 2)
    assert names(g.next()) == ["4"]
    done(g, 4)
    assert g.status == H.G_COMPLETED


def test_error_in_task_fails_graph_and_reports_running():
    g = build(3, [(1, 2), (1, 3)])
    assert names(g.next()) == ["1"]
    done(g, 1)
    assert names(g.next()) == ["2", "3"]
    g.task_done(1, ok=False, error="boom")
    assert g.status == H.G_FAILED
    failed, desc, task = g.error()
    assert failed and desc == "boom" and task == 1
    assert g.next() == []


def test_stop_kills_running():
    g = build(3, [(1, 2), (1, 3)])
    g.next()
    done(g, 1)
    g.next()
    assert names(g.stop("Stopped from test")) == ["2", "3"]
    assert g.status == H.G_FAILED
    assert [g.task_status(i) for i in (1, 2)] == [H.T_ERROR, H.T_ERROR]
    assert g.stop("again") == []


def test_cycles_run_as_one_group():
    g = build(5, [(1, 2), (2, 1), (1, 3), (3, 4), (4, 5), (5, 3)])
    assert g.has_cycle
    assert names(g.next()) == ["1", "2"]
    done(g, 2)
    assert g.next() == [] and names(g.executing()) == ["1", "2"]
    done(g, 1)
    assert names(g.next()) == ["3", "4", "5"]
    done(g, 3, 4, 5)
    assert g.status == H.G_COMPLETED
    with pytest.raises(GrpcStatusError) as e:  # graph-executor-2 refuses cycles
        GraphDispatcher(2, [(0, 1), (1, 0)], allow_cycles=False)
    assert e.value.status == "INVALID_ARGUMENT"


def test_one_vertex_and_empty():
    g = build(1, [])
    assert names(g.next()) == ["1"]
    done(g, 1)
    assert g.status == H.G_COMPLETED
    assert GraphDispatcher(0, []).status == H.G_COMPLETED


def test_bad_edge_is_validation_error():
    with pytest.raises(GrpcStatusError) as e:
        GraphDispatcher(2, [(0, 5)])
    assert e.value.status == "INVALID_ARGUMENT"


def test_topo_order_is_a_valid_sequential_order():
    g = build(10, SIMPLE_EDGES)
    order = g.topo_order()
    pos = {t: i for i, t in enumerate(order)}
    assert sorted(order) == list(range(10))
    assert all(pos[a - 1] < pos[b - 1] for a, b in SIMPLE_EDGES)


@pytest.mark.parametrize("seed", range(40))
def test_lockstep_with_bfs_oracle(seed):
    rnd = random.Random(seed)
    n = rnd.randint(1, 24)
    edges = []
    for b in range(n):
        for a in range(b):
            if rnd.random() < 0.18:
                edges.append((a, b))
    if seed % 4 == 0 and n > 3:  # sprinkle cycles
        for _ in range(rnd.randint(1, 3)):
            a, b = rnd.sample(range(n), 2)
            edges.append((a, b))
    g, o = GraphDispatcher(n, edges), BfsOracle(n, edges)
    for _ in range(4 * n + 10):
        sent = o.step()
        started = set(g.next())
        assert started == sent
        if o.status in ("COMPLETED", "FAILED"):  # the reference leaves currentExecutionGroup stale on completion
            break
        assert set(g.executing()) == o.executing_now()
        running = sorted(t for t in o.executions if o.task_status[t] == "EXECUTING")
        if not running:
            continue
        k = rnd.randint(1, len(running))
        for t in rnd.sample(running, k):
            o.change_status(t, T_SUCCESS)
            g.task_done(t)
    assert o.status == "COMPLETED" and g.status == H.G_COMPLETED


def test_every_order_of_the_reference_local_runtime_is_a_valid_dispatch_order():
    """the reference's local runtime runs the calls of a barrier one at a time in a DFS-topological order
    (tests/golden/reference_order_vectors.json, produced by its own LocalRuntime.exec); the C++ dispatcher's rule — a
    task may start once the producers of all its inputs have finished — must admit each of those sequences"""
    import importlib.util
    import json
    from pathlib import Path

    golden = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("order_battery", golden / "order_battery.py")
    battery = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(battery)
    orders = json.loads((golden / "reference_order_vectors.json").read_text())
    for name, calls in battery.graphs().items():
        index = {c.id: i for i, c in enumerate(calls)}
        producer = {eid: i for i, c in enumerate(calls) for eid in c.entry_ids}
        edges = sorted({(producer[e], i) for i, c in enumerate(calls)
                        for e in list(c.arg_entry_ids) + list(c.kwarg_entry_ids.values()) if e in producer and producer[e] != i})
        g = H.GraphDispatcher(len(calls), edges, allow_cycles=False)
        try:
            ready = set(g.next())
            for call_id in orders[name]:
                task = index[call_id]
                assert task in ready, f"{name}: {call_id} ran before its producers"
                ready.discard(task)
                g.task_done(task, True)
                ready |= set(g.next())
            assert not ready and g.status == H.G_COMPLETED
        finally:
            g.close()
