"""How a lazy proxy of a snapshot entry gets its value, observed twice: by tests/golden/make_golden.py with the
*reference's own* ``lzy_proxy`` (pylzy/lzy/api/v1/utils/proxy_adapter.py:55-83 on lzy/proxy/automagic.py and
lzy/utils/event_loop.py, loaded by file path in the authoring container) → tests/golden/reference_materialize_vectors.json,
and by tests/test_lazy_proxy.py with lzy_b200.api.v1.utils.proxy_adapter's: the supplied value wins; else what the
snapshot has; else exactly one barrier and what the snapshot has then; else the snapshot's own reason is raised, or a
RuntimeError naming the entry and the workflow.  Asking the markers never computes anything.

`ns`: ``lzy_proxy``, ``materialize``, ``materialized``, ``is_lzy_proxy``, ``get_proxy_entry_id``, ``Result``, ``Absence``."""
from typing import Any, Dict


class FakeSnapshot:
    def __init__(self, log, answers):
        self.log, self.answers = log, list(answers)

    async def get_data(self, entry_id):
        self.log.append(f"get_data({entry_id})")
        return self.answers.pop(0)


class FakeWorkflow:
    name = "wf-7"

    def __init__(self, log, answers):
        self.log, self.snapshot = log, FakeSnapshot(log, answers)

    async def _barrier(self):
        self.log.append("barrier")


def run(ns: Dict[str, Any]) -> Dict[str, Any]:
    Result, Absence = ns["Result"], ns["Absence"]
    out: Dict[str, Any] = {}

    def scenario(tag, answers, value=None):
        log: list = []
        wf = FakeWorkflow(log, answers)
        p = ns["lzy_proxy"]("entry-1", (int,), wf, value) if value is not None else ns["lzy_proxy"]("entry-1", (int,), wf)
        record: Dict[str, Any] = {"is proxy": ns["is_lzy_proxy"](p), "entry id": ns["get_proxy_entry_id"](p),
                                  "materialized before": ns["materialized"](p), "log before": list(log)}
        try:
            record["value"] = repr(ns["materialize"](p))
        except Exception as e:  # noqa: BLE001
            record["raised"] = [type(e).__name__, str(e)]
        record["log"] = list(log)
        record["materialized after"] = ns["materialized"](p) if "value" in record else None
        if "value" in record:
            record["second touch"] = str(p)  # any use of the proxy after the first: served from the cached origin
            record["log after second touch"] = list(log)
        out[tag] = record

    scenario("value supplied", [], value=Result(5))
    scenario("already in the snapshot", [Result(6)])
    scenario("after one barrier", [Absence(), Result(7)])
    scenario("absent with a cause", [Absence(), Absence(ValueError("Cannot find blob in storage by uri=s3://x"))])
    scenario("absent without a cause", [Absence(), Absence()])
    scenario("a None result is a result", [Result(None)])
    out["not a proxy"] = {"is proxy": ns["is_lzy_proxy"](5)}
    try:
        ns["get_proxy_entry_id"](5)
    except ValueError as e:
        out["not a proxy"]["entry id"] = str(e)
    return out
