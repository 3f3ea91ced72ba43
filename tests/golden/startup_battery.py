"""The op-side entry point over real files, run twice: by tests/golden/make_golden.py with the *reference's own*
``process_execution`` (pylzy/lzy/api/v1/startup.py:109-170, loaded by file path in the authoring container; it reads
and writes slot files under ``LZY_MOUNT``) to produce tests/golden/reference_startup_vectors.json, and by
tests/test_startup.py with lzy_b200.api.v1.startup's over a directory mount.  Both sides get the same serializer
registry (this package's — serialzy is absent), so what is compared is the control flow: which slots are written in
which scenario, what the exception slot holds, what is raised, what happens to ``LZY_OP_MAIN_PID``.

`ns["execute"](mount_dir, registry, op, args_paths, kwargs_paths, output_paths, exception_path, lazy_arguments)`"""
import logging
import os
from io import BytesIO
from typing import Any, Dict, List

MAIN_PID_ENV_VAR = "LZY_OP_MAIN_PID"


def op_concat(a: int, b: str) -> str:
    return f"{b}{a}"


def op_pair(a: int) -> (int, str):
    return a + 1, str(a)


def op_raises(a: int) -> int:
    raise ValueError("boom", a)


def op_wrong_type(a: int) -> int:
    return "not an int"


def op_none(a: int) -> None:
    return None


def op_lazy(a: int, b: str) -> str:
    return f"{type(a) is int}/{type(b) is str}/{a + 1}/{b}"


def op_returns_its_argument(a: List[int], k: int) -> List[int]:
    return a


def op_ignores_argument(a: int) -> int:
    return 5


def run(ns: Dict[str, Any], registry, root: str) -> Dict[str, Any]:
    logger = logging.getLogger("startup-battery")
    out: Dict[str, Any] = {}

    def scenario(tag, op, args, kwargs, outputs, lazy=False, main_pid=None, corrupt=None):
        mount = os.path.join(root, tag)
        os.makedirs(mount)

        def put(path, typ, value):
            with open(mount + path, "wb") as f:
                if corrupt == path:
                    f.write(b"this is not a pickle")
                else:
                    registry.find_serializer_by_type(typ).serialize(value, f)
            return typ, path

        args_paths = [put(f"/{tag}.a{i}", typ, v) for i, (typ, v) in enumerate(args)]
        kwargs_paths = {k: put(f"/{tag}.k.{k}", typ, v) for k, (typ, v) in kwargs.items()}
        output_paths = [(typ, f"/{tag}.r{i}") for i, typ in enumerate(outputs)]
        exception_path = (tuple, f"/{tag}.exc")
        for _typ, path in output_paths + [exception_path]:
            open(mount + path, "wb").close()  # slots exist (empty) before the op runs, as the worker leaves them
        os.environ.pop(MAIN_PID_ENV_VAR, None)
        if main_pid is not None:
            os.environ[MAIN_PID_ENV_VAR] = str(main_pid)
        raised = None
        try:
            ns["execute"](mount, registry, op, args_paths, kwargs_paths, output_paths, exception_path, lazy, logger)
        except Exception as e:  # noqa: BLE001
            raised = [type(e).__name__, [repr(a) for a in e.args]]
        record: Dict[str, Any] = {"raised": raised, "main_pid_left_in_env": MAIN_PID_ENV_VAR in os.environ}
        os.environ.pop(MAIN_PID_ENV_VAR, None)

        def slot(typ, path):
            raw = open(mount + path, "rb").read()
            if not raw:
                return "empty"
            value = registry.find_serializer_by_type(typ).deserialize(BytesIO(raw))
            if isinstance(value, tuple) and len(value) == 3 and isinstance(value[0], type) and issubclass(value[0], BaseException):
                return {"exc_info": [value[0].__name__, [repr(a) for a in value[1].args], value[2] is not None]}
            return repr(value)

        record["outputs"] = [slot(typ, path) for typ, path in output_paths]
        record["exception_slot"] = slot(*exception_path)
        out[tag] = record

    scenario("concat", op_concat, [(int, 7), (str, "x")], {}, [str])
    scenario("kwargs", op_concat, [(int, 7)], {"b": (str, "kw")}, [str])
    scenario("pair", op_pair, [(int, 1)], {}, [int, str])
    scenario("raises", op_raises, [(int, 3)], {}, [int])
    scenario("wrong-type", op_wrong_type, [(int, 3)], {}, [int])
    scenario("none", op_none, [(int, 3)], {}, [type(None)])
    scenario("lazy", op_lazy, [(int, 1), (str, "s")], {}, [str], lazy=True)
    scenario("eager", op_lazy, [(int, 1), (str, "s")], {}, [str], lazy=False)
    scenario("lazy-returns-argument", op_returns_its_argument, [(List[int], [1, 2, 3])], {"k": (int, 0)}, [List[int]], lazy=True)
    scenario("not-main-process", op_concat, [(int, 7), (str, "x")], {}, [str], main_pid=os.getpid() + 1)
    scenario("corrupt-input-eager", op_ignores_argument, [(int, 1)], {}, [int], corrupt="/corrupt-input-eager.a0")
    scenario("corrupt-input-lazy-untouched", op_ignores_argument, [(int, 1)], {}, [int], lazy=True, corrupt="/corrupt-input-lazy-untouched.a0")
    scenario("raises-lazy", op_raises, [(int, 3)], {}, [int], lazy=True)
    return out
