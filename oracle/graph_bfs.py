"""oracle/graph_bfs.py — TEST INFRASTRUCTURE.  Literal restatement of the reference's BFS dispatch:

  Algorithms.condenseGraph / findRoots / getNextBfsGroup
      lzy/graph-executor/src/main/java/ai/lzy/graph/algo/Algorithms.java:7-148
  BfsGraphProcessor.getNextExecutionGroup / nextStep / exec
      lzy/graph-executor/src/main/java/ai/lzy/graph/exec/BfsGraphProcessor.java:44-87,124-167,176-258
  DirectChannelChecker.ready
      lzy/graph-executor/src/main/java/ai/lzy/graph/exec/impl/DirectChannelChecker.java:24-32

The scheduler is replaced by a status table the test drives (SchedulerApiMock in
lzy/graph-executor/src/test/java/ai/lzy/graph/test/mocks/).  Sleeps are omitted."""
from __future__ import annotations

import sys
from typing import Dict, FrozenSet, List, Sequence, Set, Tuple

WAITING, EXECUTING, COMPLETED, FAILED = "WAITING", "EXECUTING", "COMPLETED", "FAILED"
T_NONE, T_EXECUTING, T_SUCCESS, T_ERROR = "NONE", "EXECUTING", "SUCCESS", "ERROR"


class BfsOracle:
    def __init__(self, n: int, edges: Sequence[Tuple[int, int]]):
        self.n = n
        self.children: Dict[int, List[int]] = {v: [] for v in range(n)}
        self.parents: Dict[int, List[int]] = {v: [] for v in range(n)}
        for a, b in edges:
            self.children[a].append(b)
            self.parents[b].append(a)
        self.status = WAITING
        self.task_status: Dict[int, str] = {}      # tasks sent to the scheduler
        self.executions: List[int] = []            # every task ever sent
        self.current_group: List[int] = []         # currentExecutionGroup
        self._condense()

    # -- Algorithms.condenseGraph (Kosaraju) ------------------------------------------------------
    def _condense(self):
        sys.setrecursionlimit(max(10000, 4 * self.n + 100))
        used: Set[int] = set()
        order: List[int] = []

        def top_sort(v):
            used.add(v)
            for w in self.children[v]:
                if w not in used:
                    top_sort(w)
            order.append(v)

        for v in range(self.n):
            if v not in used:
                top_sort(v)
        used.clear()
        order.reverse()
        comps: List[Set[int]] = []

        def strong(v, comp):
            used.add(v)
            comp.add(v)
            for w in self.parents[v]:
                if w not in used:
                    strong(w, comp)

        for v in order:
            if v not in used:
                c: Set[int] = set()
                strong(v, c)
                comps.append(c)
        self.comps: List[FrozenSet[int]] = [frozenset(c) for c in comps]
        self.comp_of = {v: i for i, c in enumerate(self.comps) for v in c}
        self.c_children: Dict[int, Set[int]] = {i: set() for i in range(len(comps))}
        self.c_parents: Dict[int, Set[int]] = {i: set() for i in range(len(comps))}
        self.c_parent_edges: Dict[int, List[Tuple[int, int]]] = {i: [] for i in range(len(comps))}
        for a in range(self.n):
            for b in self.children[a]:
                ca, cb = self.comp_of[a], self.comp_of[b]
                if ca != cb:
                    self.c_children[ca].add(cb)
                    self.c_parents[cb].add(ca)
                    self.c_parent_edges[cb].append((a, b))

    def _roots(self) -> Set[int]:  # Algorithms.findRoots on the condensed graph
        return {c for c in range(len(self.comps)) if not self.c_parents[c]}

    def _success(self, task: int) -> bool:
        return self.task_status.get(task) == T_SUCCESS

    # -- BfsGraphProcessor.getNextExecutionGroup ---------------------------------------------------
    def _next_group(self) -> Set[int]:
        if self.status == WAITING:
            return {v for c in self._roots() for v in self.comps[c]}
        assert self.status == EXECUTING
        current = {self.comp_of[t] for t in self.current_group}
        already = {self.comp_of[t] for t in self.executions}
        nxt: Set[int] = set()
        for comp in current:
            if not all(v in self.task_status and self._success(v) for v in self.comps[comp]):
                nxt.add(comp)
        bfs = set()
        for comp in current:  # Algorithms.getNextBfsGroup
            bfs |= self.c_children[comp]
        for comp in bfs:
            ready = all(a in self.task_status and self._success(a) for a, _b in self.c_parent_edges[comp])
            if ready and comp not in already:
                nxt.add(comp)
        return {v for c in nxt for v in self.comps[c]}

    # -- BfsGraphProcessor.nextStep / exec ------------------------------------------------------------
    def step(self) -> Set[int]:
        """one QueueManager iteration; returns the tasks newly sent to the scheduler"""
        if self.status in (COMPLETED, FAILED):
            return set()
        if self.status == EXECUTING:
            completed = 0
            for t in self.executions:
                st = self.task_status[t]
                if st == T_SUCCESS:
                    completed += 1
                if st == T_ERROR:
                    self.status = FAILED
                    for u in self.executions:  # api.kill → mock reports ERROR
                        self.task_status[u] = T_ERROR
                    return set()
            if completed == self.n:
                self.status = COMPLETED
                return set()
        new_group = self._next_group()
        diff = new_group - set(self.current_group)
        for t in diff:
            self.task_status[t] = T_EXECUTING
            self.executions.append(t)
        self.current_group = [t for t in self.executions if t in new_group]
        self.status = EXECUTING
        if self.n == 0:
            self.status = COMPLETED
        return diff

    def change_status(self, task: int, st: str):
        self.task_status[task] = st

    def executing_now(self) -> Set[int]:
        return set(self.current_group)
