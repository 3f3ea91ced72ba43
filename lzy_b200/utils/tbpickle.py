"""Makes ``sys.exc_info()`` tuples picklable: the exception slot of an @op carries
``(type, value, traceback)`` (pylzy/lzy/api/v1/startup.py:137-170) and the caller re-raises
``value.with_traceback(tb)`` (pylzy/lzy/api/v1/local/runtime.py:165-173).  A traceback is flattened to
(file, function, line) records and rebuilt as a chain of synthetic frames, which is enough for
``traceback.extract_tb`` to report the original locations (pylzy/tests/api/v1/test_startup.py:52-83)."""
from __future__ import annotations

import copyreg
import sys
import types
from typing import List, Tuple

Record = Tuple[str, str, int]


class _Probe(Exception):
    pass


def flatten(tb) -> List[Record]:
    out: List[Record] = []
    while tb is not None:
        code = tb.tb_frame.f_code
        out.append((code.co_filename, code.co_name, tb.tb_lineno))
        tb = tb.tb_next
    return out


def _frame_at(filename: str, name: str, lineno: int):
    src = "\n" * max(0, lineno - 1) + "raise __probe__"
    code = compile(src, filename, "exec").replace(co_name=name)
    try:
        exec(code, {"__probe__": _Probe()})
    except _Probe:
        return sys.exc_info()[2].tb_next.tb_frame  # type: ignore[union-attr]
    raise AssertionError("unreachable")


def rebuild(records: List[Record]):
    tb = None
    for filename, name, lineno in reversed(records):
        frame = _frame_at(filename, name, lineno)
        tb = types.TracebackType(tb, frame, frame.f_lasti, lineno)
    return tb


def _reduce(tb):
    return rebuild, (flatten(tb),)


def install() -> None:
    copyreg.pickle(types.TracebackType, _reduce)
