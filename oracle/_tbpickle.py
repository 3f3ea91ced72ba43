"""oracle/_tbpickle.py — TEST INFRASTRUCTURE.  Makes ``sys.exc_info()`` tuples picklable the way the
reference's serializer stack does for the exception slot (pylzy/lzy/api/v1/startup.py:137-170 writes
``sys.exc_info()``; pylzy/lzy/api/v1/local/runtime.py:165-173 re-raises ``value.with_traceback(tb)``;
pylzy/tests/api/v1/test_startup.py:52-83 checks ``traceback.extract_tb`` of the loaded traceback).
A traceback travels as a list of (filename, function name, line number) and is rebuilt from
synthetic frames."""
import copyreg
import sys
import types


def _flatten(tb):
    out = []
    while tb is not None:
        co = tb.tb_frame.f_code
        out.append((co.co_filename, co.co_name, tb.tb_lineno))
        tb = tb.tb_next
    return out


def _fake_frame(filename, name, lineno):
    code = compile("\n" * (lineno - 1) + "raise __tb_marker__", filename, "exec")
    code = code.replace(co_name=name)
    try:
        exec(code, {"__tb_marker__": _Marker()})
    except _Marker:
        return sys.exc_info()[2].tb_next.tb_frame
    raise RuntimeError("unreachable")


class _Marker(Exception):
    pass


def _rebuild(frames):
    nxt = None
    for filename, name, lineno in reversed(frames):
        fr = _fake_frame(filename, name, lineno)
        nxt = types.TracebackType(nxt, fr, fr.f_lasti, lineno)
    return nxt


def _reduce_tb(tb):
    return _rebuild, (_flatten(tb),)


def install():
    copyreg.pickle(types.TracebackType, _reduce_tb)
