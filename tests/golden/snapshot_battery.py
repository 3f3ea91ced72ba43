"""The id → data table of a workflow, exercised twice over a real directory: by tests/golden/make_golden.py with the
*reference's own* ``DefaultSnapshot`` on its own ``FsStorageClient`` (pylzy/lzy/api/v1/snapshot.py:77-191,
pylzy/lzy/storage/async_/fs.py, loaded by file path in the authoring container; serializers: this package's
registry, serialzy being absent) to produce tests/golden/reference_snapshot_vectors.json, and by
tests/test_abi_and_oracle.py with the oracle's restatement (``oracle/local_runtime.py: OracleSnapshot``), which is
what the CPU baseline and the local-path parity tests run on.  Content addressing (URI = prefix + md5 of the
serialized bytes), equal data stored once, and what reading an entry answers when there is nothing to read.

`ns`: ``create(name, typ) -> entry``, ``put(entry, value)``, ``get(entry) -> ("result", value) | ("absence", message)``,
``uri(entry)`` / ``hash(entry)`` (None when unset), ``set_uri(entry, uri)``, ``stored() -> sorted file names``."""
from typing import Any, Dict, List


def run(ns: Dict[str, Any], prefix: str) -> Dict[str, Any]:
    out: Dict[str, Any] = {}

    def read(entry):
        answer = list(ns["get"](entry))
        return answer if answer[0] == "result" else ["absence"]

    rel = lambda u: None if u is None else (u[len(prefix):] if u.startswith(prefix) else u)  # noqa: E731

    a = ns["create"]("f.a", int)
    out["fresh"] = {"uri": ns["uri"](a), "hash": ns["hash"](a), "get": read(a)}
    ns["put"](a, 42)
    out["after put"] = {"uri": rel(ns["uri"](a)), "hash": ns["hash"](a), "stored": ns["stored"](), "get": read(a)}

    b = ns["create"]("f.b", int)
    ns["put"](b, 42)
    out["equal data"] = {"same uri": ns["uri"](b) == ns["uri"](a), "stored": len(ns["stored"]())}

    c = ns["create"]("g.c", str)
    ns["put"](c, "forty-two")
    d = ns["create"]("g.d", List[int])
    ns["put"](d, [4, 2])
    out["other data"] = {"uris": [rel(ns["uri"](c)), rel(ns["uri"](d))], "stored": len(ns["stored"]()),
                         "get": [read(c), read(d)]}

    e = ns["create"]("h.result", int)
    ns["set_uri"](e, prefix + "../ops/h/call-1/return_0")
    out["assigned but not produced"] = {"get": read(e)}
    ns["put"](e, 7)  # put re-addresses by content
    out["put overrides an assigned uri"] = {"uri": rel(ns["uri"](e)), "get": read(e)}

    ns["put"](a, 43)  # an entry may be written again: it moves to the new content's address
    out["rewritten"] = {"uri": rel(ns["uri"](a)), "stored": len(ns["stored"]()), "b still reads": read(b)}
    return out
