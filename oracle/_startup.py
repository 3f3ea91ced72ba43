"""oracle/_startup.py — TEST INFRASTRUCTURE: the op trampoline of the reference local runtime,
restated (pylzy/lzy/api/v1/startup.py:31-49 read_data, :52-77 write_data, :109-170
process_execution, :185-224 main).  Run as `python -u _startup.py <base64 cloudpickle request>`
with LZY_MOUNT pointing at the per-call directory.  cloudpickle stands in for serialzy
(parity unpinned at the byte level, see oracle/__init__.py)."""
import base64
import os
import sys

import cloudpickle

MAIN_PID_ENV_VAR = "LZY_OP_MAIN_PID"
_read_cache = {}


def read_data(path):
    if path in _read_cache:
        return _read_cache[path]
    mount = os.environ["LZY_MOUNT"]
    with open(mount + path, "rb") as f:
        data = cloudpickle.load(f)
    _read_cache[path] = data
    return data


def write_data(path, data):
    mount = os.environ["LZY_MOUNT"]
    with open(mount + path, "wb") as out:
        cloudpickle.dump(data, out)
        out.flush()


class _Lazy:
    """lazy_arguments=True hands the op a proxy that reads on first touch (startup.py:127-136)."""

    def __init__(self, path):
        object.__setattr__(self, "_p", path)

    def __getattr__(self, name):
        return getattr(read_data(object.__getattribute__(self, "_p")), name)


def process_execution(req):
    pid = os.getpid()
    if MAIN_PID_ENV_VAR not in os.environ:
        os.environ[MAIN_PID_ENV_VAR] = str(pid)
    main_pid = int(os.environ[MAIN_PID_ENV_VAR])
    exc_path = req["exception_path"][1]
    try:
        args = [read_data(p) for _t, p in req["args_paths"]]
        kwargs = {n: read_data(p) for n, (_t, p) in req["kwargs_paths"].items()}
    except Exception:
        write_data(exc_path, sys.exc_info())
        raise
    try:
        res = req["op"](*args, **kwargs)
    except Exception:
        write_data(exc_path, sys.exc_info())
        raise
    finally:
        del os.environ[MAIN_PID_ENV_VAR]
    if main_pid != pid:
        return
    outs = req["output_paths"]
    try:
        if len(outs) == 1:
            write_data(outs[0][1], res)
        else:
            for (_t, p), d in zip(outs, res):
                write_data(p, d)
        write_data(exc_path, None)
    except Exception:
        write_data(exc_path, sys.exc_info())
        raise


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import _tbpickle
    _tbpickle.install()
    process_execution(cloudpickle.loads(base64.b64decode(sys.argv[1].encode("ascii"))))
