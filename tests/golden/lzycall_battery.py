"""What recording one ``@op`` call does to the snapshot, observed twice: by tests/golden/make_golden.py with the
*reference's own* ``LzyCall`` (pylzy/lzy/core/call.py:38-120, loaded by file path in the authoring container) to produce
tests/golden/reference_lzycall_vectors.json, and by tests/test_call_signatures.py with lzy_b200.core.call's.  Compared:
the entries created (name, type, in order), which values are written to the snapshot right away — local arguments,
at construction, before anything runs — and which arguments reuse the entry of the op that produced them, and the
combined environment (Lzy, then workflow, then call).

`ns`: ``bind`` (argument binding of that side), ``LzyCall``, ``lzy_proxy``, ``Result``, ``registry``."""
from types import SimpleNamespace
from typing import Any, Dict, List, Optional


def f_mixed(a: int, b: str, *rest, k: float = 0.5, **named) -> (int, str):
    ...


def f_none(x: Optional[int]) -> None:
    ...


def f_list(items: List[int]) -> List[int]:
    ...


class Env:
    """stands for LzyEnvironment: combine() keeps the order in which levels were applied"""

    def __init__(self, levels):
        self.levels = list(levels)

    def combine(self, other):
        return Env(self.levels + other.levels)

    def validate(self):
        self.levels.append("validated")


def _type_name(t) -> str:
    return getattr(t, "__name__", None) if isinstance(t, type) else repr(t).replace("typing.", "")


class RecordingSnapshot:
    def __init__(self, log):
        self.log, self.entries, self.n = log, {}, 0

    def create_entry(self, name, typ):
        self.n += 1
        entry = SimpleNamespace(id=f"entry-{self.n}", name=name, typ=typ)
        self.entries[entry.id] = entry
        self.log.append(["create_entry", name, _type_name(typ)])
        return entry

    def get(self, eid):
        return self.entries[eid]

    async def put_data(self, eid, data):
        self.log.append(["put_data", self.entries[eid].name, repr(data)])


def run(ns: Dict[str, Any]) -> Dict[str, Any]:
    out: Dict[str, Any] = {}

    def record(tag, f, args, kwargs, output_type, proxies=(), cache=False):
        log: list = []
        snapshot = RecordingSnapshot(log)
        workflow = SimpleNamespace(owner=SimpleNamespace(env=Env(["lzy"]), serializer_registry=ns["registry"]),
                                   env=Env(["workflow"]), snapshot=snapshot, name="wf")
        for eid, typ, value in proxies:  # results of earlier ops: their entries exist already
            snapshot.entries[eid] = SimpleNamespace(id=eid, name=eid, typ=typ)
        made = {eid: ns["lzy_proxy"](eid, (typ,), workflow, ns["Result"](value)) for eid, typ, value in proxies}
        args = [made.get(a, a) if isinstance(a, str) else a for a in args]
        kwargs = {k: (made.get(v, v) if isinstance(v, str) else v) for k, v in kwargs.items()}
        signature = ns["bind"](f, *args, output_type=output_type, snapshot=snapshot,
                               serializer_registry=ns["registry"], **kwargs)
        call = ns["LzyCall"](workflow=workflow, signature=signature, call_env=Env(["call"]), description="d",
                             version="1.0", cache=cache, lazy_arguments=False)
        n_declared = f.__code__.co_argcount
        made_up = list(signature.func.arg_names[n_declared:])
        for row in log:  # normalise the made-up names of *args
            for i, name in enumerate(made_up):
                row[1] = row[1].replace(name, f"*{i}")
        out[tag] = {"log": log, "arg_entry_ids": list(call.arg_entry_ids), "kwarg_entry_ids": dict(call.kwarg_entry_ids),
                    "entry_ids": list(call.entry_ids), "exception_id": call.exception_id, "callable_name": call.callable_name,
                    "env": call.final_env.levels, "cache": call.cache, "version": call.version}

    record("mixed", f_mixed, [1, "s", 2.5, None], {"k": 1.5, "extra": [1]}, (int, str))
    record("with-proxies", f_mixed, ["p-int", "p-str", "p-int"], {"extra": "p-str"}, (int, str),
           proxies=[("p-int", int, 5), ("p-str", str, "five")])
    record("none", f_none, [None], {}, (type(None),))
    record("list", f_list, [[1, 2, 3]], {}, (List[int],), cache=True)
    record("same-value-twice", f_mixed, [7, "x", 7, 7], {}, (int, str))
    return out
