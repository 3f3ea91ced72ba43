"""``@op`` and calling the decorated function, observed twice: by tests/golden/make_golden.py with the *reference's own*
decorator and ``LazyCallWrapper`` (pylzy/lzy/core/op.py:18-61, pylzy/lzy/core/call.py:189-250, loaded by file path in
the authoring container) → tests/golden/reference_wrapper_vectors.json, and by tests/test_api_core.py with
lzy_b200.core's.  What the decorator records (result types from the annotation, options, the function's name and
docstring), the refusal of an unannotated function, the plain call outside a workflow, and inside one: the call is
registered, one lazy proxy per result bound to the call's result entries, the real ``None`` for ``-> None``, values
instead of proxies when the workflow is eager.

`ns`: ``op``, ``LzyWorkflow`` (that side's class: the battery sets / clears its active instance), ``is_lzy_proxy``,
``get_proxy_entry_id``, ``Result``, ``Env`` (a stand-in environment class accepted by that side)."""
from types import SimpleNamespace
from typing import Any, Dict, List, Optional, Tuple


def _type_name(t) -> str:
    return getattr(t, "__name__", None) if isinstance(t, type) else repr(t).replace("typing.", "")


class Snapshot:
    def __init__(self, Result):
        self.entries, self.n, self.Result = {}, 0, Result

    def create_entry(self, name, typ):
        self.n += 1
        entry = SimpleNamespace(id=f"entry-{self.n}", name=name, typ=typ)
        self.entries[entry.id] = entry
        return entry

    def get(self, eid):
        return self.entries[eid]

    async def put_data(self, eid, data):
        pass

    async def get_data(self, eid):  # an eager workflow has run the call: every result is there
        return self.Result(f"value of {self.entries[eid].name}")


def run(ns: Dict[str, Any]) -> Dict[str, Any]:
    op, out = ns["op"], {}

    def plain(a: int, b: int = 2) -> int:
        """adds"""
        return a + b

    def pair(a: int) -> (int, str):
        return a, str(a)

    def typing_pair(a: int) -> Tuple[int, str]:
        return a, str(a)

    def nothing(a: int) -> None:
        return None

    def maybe(a: int) -> Optional[List[int]]:
        return None

    def unannotated(a):
        return a

    decorated = {"plain": op(plain), "pair": op(pair), "typing_pair": op(typing_pair), "nothing": op(nothing), "maybe": op(maybe),
                 "with options": op(cache=True, version="3.1", description="d", lazy_arguments=True)(plain),
                 "explicit outputs": op(output_types=(str,))(unannotated)}
    out["decorated"] = {k: {"output_types": [_type_name(t) for t in w.output_types], "cache": w.cache, "version": w.version,
                            "description": w.description, "lazy_arguments": w.lazy_arguments, "name": w.__name__,
                            "doc": w.__doc__} for k, w in decorated.items()}
    try:
        op(unannotated)
        out["unannotated"] = "accepted"
    except TypeError as e:
        text = str(e)
        out["unannotated"] = text[:text.index(" at ")] + " at <file>:<line>" + text[text.index(". Please"):]

    out["outside a workflow"] = {"plain(1)": decorated["plain"](1), "pair(3)": list(decorated["pair"](3)), "nothing(1)": decorated["nothing"](1)}

    def inside(eager):
        registered = []
        snapshot = Snapshot(ns["Result"])
        wf = SimpleNamespace(name="wf", eager=eager, snapshot=snapshot, env=ns["Env"](), register_call=registered.append,
                             owner=SimpleNamespace(env=ns["Env"](), serializer_registry=ns["registry"]))
        ns["LzyWorkflow"].instance = wf
        try:
            record = {}
            for label, call in {"plain(1)": lambda: decorated["plain"](1), "plain(1, b=5)": lambda: decorated["plain"](1, b=5),
                                "pair(3)": lambda: decorated["pair"](3), "typing_pair(3)": lambda: decorated["typing_pair"](3),
                                "nothing(1)": lambda: decorated["nothing"](1), "maybe(1)": lambda: decorated["maybe"](1)}.items():
                before = len(registered)
                answer = call()
                recorded = registered[before:]
                answers = list(answer) if isinstance(answer, tuple) else [answer]
                record[label] = {
                    "registered": len(recorded), "callable": recorded[0].callable_name, "results": len(recorded[0].entry_ids),
                    "answer": ["proxy of " + snapshot.get(ns["get_proxy_entry_id"](a)).name if ns["is_lzy_proxy"](a) else repr(a)
                               for a in answers],
                    "a tuple": isinstance(answer, tuple),
                    "bound to the call's entries": [ns["get_proxy_entry_id"](a) for a in answers if ns["is_lzy_proxy"](a)] in
                                                   ([], list(recorded[0].entry_ids))}
            return record
        finally:
            ns["LzyWorkflow"].instance = None

    out["inside a lazy workflow"] = inside(False)
    out["inside an eager workflow"] = inside(True)
    return out
