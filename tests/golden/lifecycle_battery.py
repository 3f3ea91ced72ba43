"""Entering, using and leaving ``with lzy.workflow(...)``, observed twice: by tests/golden/make_golden.py with the
*reference's own* ``LzyWorkflow`` (pylzy/lzy/core/workflow.py:118-236, loaded by file path in the authoring container)
driving a recording runtime → tests/golden/reference_lifecycle_vectors.json, and by tests/test_api_core.py with
lzy_b200.core.workflow's.  Recorded per scenario: the runtime calls in order (``start``, ``exec(n)``, ``finish``,
``abort``), what the block raised, which workflow is active afterwards, where the snapshot keeps local data.

One scenario is reported but not compared — ``second workflow while one is active``: the reference aborts the running
execution when the second ``__enter__`` is refused (its ``__abort`` runs unconditionally); this package leaves the
active workflow alone (tests/test_api_core.py::test_a_refused_second_workflow_leaves_the_active_one_running).

`ns`: ``make_workflow(name, runtime, log, eager) -> workflow`` (that side's class over that side's owner),
``active() -> name of the active workflow or None``."""
import os
from types import SimpleNamespace
from typing import Any, Dict


class RecordingRuntime:
    def __init__(self, log, fail_start=False, fail_exec=False):
        self.log, self.fail_start, self.fail_exec = log, fail_start, fail_exec

    async def start(self, workflow):
        self.log.append("start")
        if self.fail_start:
            raise ConnectionError("cannot start")
        return "execution-1"

    async def exec(self, calls, progress):
        self.log.append(f"exec({len(calls)})")
        if self.fail_exec:
            raise RuntimeError("graph failed")

    async def finish(self):
        self.log.append("finish")

    async def abort(self):
        self.log.append("abort")

    async def storage(self):
        return None


def fake_call(tag):
    def f():
        pass

    return SimpleNamespace(id=f"call-{tag}", callable_name="f", signature=SimpleNamespace(func=SimpleNamespace(callable=f)), cache=False, version="0.0", arg_entry_ids=[], kwarg_entry_ids={},
                           kwargs={}, entry_ids=[f"{tag}.r0"], exception_id=f"{tag}.exc")


def run(ns: Dict[str, Any], user) -> Dict[str, Any]:
    out: Dict[str, Any] = {"user": user}

    def scenario(tag, body, eager=False, **runtime_options):
        log: list = []
        runtime = RecordingRuntime(log, **runtime_options)
        wf = ns["make_workflow"]("wf", runtime, log, eager)
        raised = None
        try:
            with wf:
                body(wf, log)
        except Exception as e:  # noqa: BLE001
            raised = [type(e).__name__, str(e)]
        out[tag] = {"log": log, "raised": raised, "active afterwards": ns["active"]()}

    def register(n):
        def body(wf, log):
            for i in range(n):
                entries = wf.snapshot.entries
                for eid in (f"c{i}.r0", f"c{i}.exc"):
                    entries[eid] = SimpleNamespace(id=eid, storage_uri=None, data_hash=None)
                wf.register_call(fake_call(f"c{i}"))
            log.append(f"queued {len(wf.call_queue)}")
        return body

    def barrier_twice(wf, log):
        register(2)(wf, log)
        wf.barrier()
        log.append(f"queued {len(wf.call_queue)}")
        wf.barrier()  # nothing queued: the runtime is not asked
        register(1)(wf, log)

    def raises(wf, log):
        register(1)(wf, log)
        raise ZeroDivisionError("in the block")

    scenario("empty block", lambda wf, log: log.append("active inside: " + str(ns["active"]())))
    scenario("two calls, run on exit", register(2))
    scenario("explicit barriers", barrier_twice)
    scenario("eager", register(2), eager=True)
    scenario("exception in the block", raises)
    scenario("runtime cannot start", register(1), fail_start=True)
    scenario("graph fails on exit", register(1), fail_exec=True)
    scenario("graph fails at a barrier", lambda wf, log: (register(1)(wf, log), wf.barrier()), fail_exec=True)
    scenario("snapshot prefix", lambda wf, log: log.append("prefix " + wf.snapshot.prefix))
    return out
