"""Storage URIs of op results, assigned twice: by tests/golden/make_golden.py with the *reference's own* ``LzyWorkflow``
(pylzy/lzy/core/workflow.py:247-298, loaded by file path in the authoring container, its whiteboard / logging /
environment imports stubbed) to produce tests/golden/reference_uri_vectors.json, and by tests/test_api_core.py with
lzy_b200.core.workflow's.  These URIs are the on-disk / cache naming contract: equal ``cache=True``
calls must meet at equal URIs across runs and across implementations.

`ns`: ``make_workflow(name, storage_uri)`` → a workflow object of that side whose snapshot is `snapshot`;
``assign(workflow, call)`` → runs that side's URI assignment for one call."""
import hashlib
from types import SimpleNamespace
from typing import Any, Dict


class Entry:
    def __init__(self, eid, data_hash=None):
        self.id, self.data_hash, self.storage_uri = eid, data_hash, None


class Snapshot:
    def __init__(self):
        self.entries: Dict[str, Entry] = {}

    def add(self, eid, data_hash=None):
        self.entries[eid] = Entry(eid, data_hash)
        return eid

    def get(self, eid):
        return self.entries[eid]


def _call(snapshot, tag, name, cache, version, args, kwargs, n_out):
    """args / kwargs: data hashes of the inputs (what an argument entry carries once it is stored)"""
    return SimpleNamespace(
        id=f"call-{tag}", callable_name=name, cache=cache, version=version,
        arg_entry_ids=[snapshot.add(f"{tag}.a{i}", h) for i, h in enumerate(args)],
        kwarg_entry_ids={k: snapshot.add(f"{tag}.k.{k}", h) for k, h in kwargs.items()},
        kwargs={k: None for k in kwargs},
        entry_ids=[snapshot.add(f"{tag}.r{i}") for i in range(n_out)],
        exception_id=snapshot.add(f"{tag}.exc"))


def run(ns: Dict[str, Any], user) -> Dict[str, Any]:
    h = lambda s: hashlib.md5(s.encode()).hexdigest()  # noqa: E731
    out: Dict[str, Any] = {}
    for storage in ("s3://bucket/prefix", "file:///tmp/lzy_storage"):
        snapshot = Snapshot()
        wf = ns["make_workflow"]("wf-name_1", storage, snapshot)
        calls = [
            _call(snapshot, "plain", "foo", False, "0.0", [h("x")], {}, 1),
            _call(snapshot, "plain2", "foo", False, "0.0", [h("x")], {"k": h("y")}, 2),
            _call(snapshot, "cached", "foo", True, "1.0", [h("x"), h("y")], {}, 1),
            _call(snapshot, "cached-same", "foo", True, "1.0", [h("x"), h("y")], {}, 1),
            _call(snapshot, "cached-swapped", "foo", True, "1.0", [h("y"), h("x")], {}, 1),
            _call(snapshot, "cached-version", "foo", True, "1.1", [h("x"), h("y")], {}, 1),
            _call(snapshot, "cached-kwargs", "bar", True, "1.0", [h("x")], {"zeta": h("z"), "alpha": h("a"), "mid": h("m")}, 3),
            _call(snapshot, "cached-kwargs-other-order", "bar", True, "1.0", [h("x")], {"mid": h("m"), "alpha": h("a"), "zeta": h("z")}, 3),
            _call(snapshot, "cached-noargs", "baz", True, "2", [], {}, 1),
            _call(snapshot, "cached-only-kwargs", "baz", True, "2", [], {"k": h("v")}, 1),
            _call(snapshot, "cached-empty-version", "qux", True, "", [h("x")], {}, 1),
            _call(snapshot, "none", "returns_none", False, "0.0", [], {}, 1),
        ]
        for c in calls:
            ns["assign"](wf, c)
        # a second pass must not move anything (entries are filled once)
        before = {eid: (e.storage_uri, e.data_hash) for eid, e in snapshot.entries.items()}
        for c in calls:
            ns["assign"](wf, c)
        assert before == {eid: (e.storage_uri, e.data_hash) for eid, e in snapshot.entries.items()}
        out[storage] = {eid: {"uri": e.storage_uri, "hash": e.data_hash} for eid, e in snapshot.entries.items()
                        if ".r" in eid or eid.endswith(".exc")}
    names = {}
    for name in ["wf", "wf-1", "wf_1", "WF9", "", "wf 1", "wf!", "wf.1", "wf/1", "ʎzy", "wf\n", "-", "a" * 300, "wf\t", "wf\n\n"]:
        try:
            ns["make_workflow"](name, "s3://b", Snapshot())
            names[repr(name[:12])] = "accepted"
        except ValueError as e:
            names[repr(name[:12])] = str(e)
    return {"user": user, "storages": out, "workflow names": names}
