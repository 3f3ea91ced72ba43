"""The order in which the reference's local runtime executes the calls of one barrier (a DFS-based topological order,
pylzy/lzy/api/v1/local/runtime.py:44-85), observed by tests/golden/make_golden.py on the *reference's own*
``LocalRuntime.exec`` with the per-call execution replaced by a recorder, for forty seeded random graphs →
tests/golden/reference_order_vectors.json; tests/test_abi_and_oracle.py holds ``OracleLocalRuntime.topo_order`` to
it, and tests/test_graph_dispatch.py checks that every such order is also a valid order for the C++ dispatcher's
dependency rule.

`ns["order"](calls) -> list of call ids in execution order`; a call is anything with ``id``, ``arg_entry_ids``,
``kwarg_entry_ids`` (dict), ``entry_ids``."""
import random
from types import SimpleNamespace
from typing import Any, Dict, List


def graphs() -> Dict[str, List[SimpleNamespace]]:
    out = {}
    for seed in range(40):
        rnd = random.Random(seed)
        n = rnd.randint(1, 9)
        calls, available = [], [f"local-{seed}-{i}" for i in range(2)]  # entries of local data, produced by no call
        for i in range(n):
            n_args = rnd.randint(0, min(3, len(available)))
            args = rnd.sample(available, n_args)
            kwargs = {}
            if available and rnd.random() < 0.4:
                kwargs["k"] = rnd.choice(available)
            outs = [f"g{seed}c{i}r{j}" for j in range(rnd.randint(1, 2))]
            calls.append(SimpleNamespace(id=f"g{seed}c{i}", arg_entry_ids=args, kwarg_entry_ids=kwargs, entry_ids=outs,
                                         exception_id=f"g{seed}c{i}exc"))
            available += outs
        rnd.shuffle(calls) if seed % 4 == 3 else None  # the API always submits producers first; a few graphs do not
        out[f"graph-{seed}"] = calls
    return out


def run(ns: Dict[str, Any]) -> Dict[str, Any]:
    return {name: ns["order"](calls) for name, calls in graphs().items()}
