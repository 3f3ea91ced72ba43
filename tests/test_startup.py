"""The op-side entry point (`lzy_b200.api.v1.startup`) against the reference's own startup tests
(pylzy/tests/api/v1/test_startup.py:15-83: test_startup, test_startup_with_exception) and the invariants of
SURVEY Appendix A.1 that live in it (lazy arguments, result type check, None in the exception slot on
success, results from the main process only).  The mount is a dict of byte streams — the role the temp files
under ``_lzy_mount`` play in the reference's tests."""
import os
import traceback
from io import BytesIO
from tempfile import NamedTemporaryFile

import pytest

from lzy_b200.api.v1 import startup
from lzy_b200.api.v1.startup import ProcessingRequest, SlotMount
from lzy_b200.api.v1.utils.pickle import pickle, unpickle
from lzy_b200.serialization.registry import LzySerializerRegistry
from lzy_b200.types import File


class BytesMount(SlotMount):
    def __init__(self):
        self.files = {}
        self.reads = []

    def put(self, path, typ, value, ser):
        buf = BytesIO()
        ser.find_serializer_by_type(typ).serialize(value, buf)
        self.files[path] = buf.getvalue()

    def get(self, path, typ, ser):
        return ser.find_serializer_by_type(typ).deserialize(BytesIO(self.files[path]), typ if isinstance(typ, type) else None)

    def read(self, path, typ, serializers):
        self.reads.append(path)
        return self.get(path, typ, serializers)

    def write(self, path, typ, data, serializers):
        self.put(path, typ, data, serializers)


def request(op, ser, **kw):
    base = dict(logger_config={}, serializers=ser.imports(), op=op, args_paths=[], kwargs_paths={}, output_paths=[],
                exception_path=(tuple, "/c/exc"))
    base.update(kw)
    return ProcessingRequest(**base)


def test_pickled_request_round_trips_as_ascii():
    s = pickle({"a": (1, 2.5, "x")})
    assert s.isascii() and " " not in s and unpickle(s) == {"a": (1, 2.5, "x")}


def test_startup():
    def test(a: str, *, b: File) -> str:
        with b.open("r") as f:
            return a + f.readline()

    with NamedTemporaryFile() as data_file:
        with open(data_file.name, "w") as f:
            f.write("2")
        ser, mount = LzySerializerRegistry(), BytesMount()
        mount.put("/c/a", str, "4", ser)
        mount.put("/c/b", File, File(data_file.name), ser)
        req = request(test, ser, args_paths=[(str, "/c/a")], kwargs_paths={"b": (File, "/c/b")}, output_paths=[(str, "/c/ret")])
        startup.main(pickle(req), mount)
        assert mount.get("/c/ret", str, ser) == "42"
        assert mount.get("/c/exc", type(None), ser) is None  # success leaves None in the exception slot


def test_startup_with_exception():
    def test() -> None:
        raise ValueError("test")

    ser, mount = LzySerializerRegistry(), BytesMount()
    req = request(test, ser)
    original_traceback = None
    try:
        startup.main(pickle(req), mount)
    except ValueError as e:
        original_traceback = e.__traceback__
    assert original_traceback is not None
    exc = mount.get("/c/exc", tuple, ser)
    assert type(exc[1]) is ValueError and exc[1].args == ("test",)
    current = traceback.extract_tb(exc[2])
    assert traceback.extract_tb(original_traceback)[-len(current):] == current
    assert "/c/ret" not in mount.files


def test_lazy_arguments_are_read_on_first_touch_and_only_once():
    def op(a: str, b: str) -> str:
        lazy = type(a) is not str  # a proxy, not the object (tests/core/test_workflow.py:386-403)
        return f"{lazy}:{a + a}"  # b is never touched

    ser, mount = LzySerializerRegistry(), BytesMount()
    mount.put("/c/a", str, "x", ser)
    mount.put("/c/b", str, "y", ser)
    req = request(op, ser, args_paths=[(str, "/c/a"), (str, "/c/b")], output_paths=[(str, "/c/ret")], lazy_arguments=True)
    startup.main(pickle(req), mount)
    assert mount.reads == ["/c/a"] and mount.get("/c/ret", str, ser) == "True:xx"

    mount.reads.clear()
    req = request(op, ser, args_paths=[(str, "/c/a"), (str, "/c/b")], output_paths=[(str, "/c/ret")], lazy_arguments=False)
    startup.main(pickle(req), mount)
    assert mount.reads == ["/c/a", "/c/b"] and mount.get("/c/ret", str, ser) == "False:xx"


def test_returning_a_lazy_argument_writes_the_real_object():
    def ident(a: str) -> str:
        return a

    ser, mount = LzySerializerRegistry(), BytesMount()
    mount.put("/c/a", str, "payload", ser)
    req = request(ident, ser, args_paths=[(str, "/c/a")], output_paths=[(str, "/c/ret")])
    startup.main(pickle(req), mount)
    assert mount.get("/c/ret", str, ser) == "payload"


def test_wrong_result_type_goes_to_the_exception_slot():
    def bad() -> int:
        return "not an int"  # type: ignore[return-value]

    ser, mount = LzySerializerRegistry(), BytesMount()
    req = request(bad, ser, output_paths=[(int, "/c/ret")])
    with pytest.raises(TypeError, match="Invalid types"):
        startup.main(pickle(req), mount)
    exc = mount.get("/c/exc", tuple, ser)
    assert exc[0] is TypeError and "/c/ret" not in mount.files


def test_multiple_outputs_are_zipped_with_their_slots():
    def two() -> (int, str):
        return 7, "seven"

    ser, mount = LzySerializerRegistry(), BytesMount()
    req = request(two, ser, output_paths=[(int, "/c/r0"), (str, "/c/r1")])
    startup.main(pickle(req), mount)
    assert mount.get("/c/r0", int, ser) == 7 and mount.get("/c/r1", str, ser) == "seven"


def test_only_the_main_process_writes_results():
    def op() -> int:
        return 1

    ser, mount = LzySerializerRegistry(), BytesMount()
    os.environ[startup.MAIN_PID_ENV_VAR] = str(os.getpid() + 1)  # as inherited by a replica (DDP) of the op process
    try:
        startup.main(pickle(request(op, ser, output_paths=[(int, "/c/ret")])), mount)
    finally:
        os.environ.pop(startup.MAIN_PID_ENV_VAR, None)
    assert mount.files == {}
    startup.main(pickle(request(op, ser, output_paths=[(int, "/c/ret")])), mount)
    assert mount.get("/c/ret", int, ser) == 1 and startup.MAIN_PID_ENV_VAR not in os.environ


def test_main_needs_a_mount():
    with pytest.raises(AssertionError, match="no slot mount"):
        startup.main("AAAA")


def test_a_request_pickled_by_a_stock_client_resolves_to_this_package():
    """cloudpickle stores `ProcessingRequest` by reference to ``lzy.api.v1.startup``; with `install_as_lzy()` that
    name is this package's module object, so the request a stock pylzy client put on the command line runs here"""
    import importlib

    import lzy_b200
    lzy_b200.install_as_lzy()
    assert importlib.import_module("lzy.api.v1.startup") is startup  # the same module object, not a second copy

    def op(a: str) -> str:
        return a.upper()

    ser, mount = LzySerializerRegistry(), BytesMount()
    mount.put("/c/a", str, "x", ser)
    req = request(op, ser, args_paths=[(str, "/c/a")], output_paths=[(str, "/c/ret")])
    ProcessingRequest.__module__ = "lzy.api.v1.startup"  # what the class is called in a stock client's pickle
    try:
        arg = pickle(req)
    finally:
        ProcessingRequest.__module__ = "lzy_b200.api.v1.startup"
    import base64
    assert b"lzy.api.v1.startup" in base64.b64decode(arg) and b"lzy_b200.api.v1.startup" not in base64.b64decode(arg)
    startup.main(arg, mount)
    assert mount.get("/c/ret", str, ser) == "X"


def test_request_encoding_matches_the_reference_text_for_text():
    """tests/golden/reference_pickle_vectors.json: output of the reference's own pickle() (utils/pickle.py:7-12)"""
    import json
    from pathlib import Path
    for v in json.loads((Path(__file__).parent / "golden" / "reference_pickle_vectors.json").read_text()):
        obj = unpickle(v["text"])
        assert repr(obj) == v["repr"]
        assert pickle(obj) == v["text"]


def test_process_execution_behaves_as_the_reference_entry_point_does(tmp_path):
    """tests/golden/reference_startup_vectors.json: thirteen scenarios run through pylzy/lzy/api/v1/startup.py's own
    process_execution over slot files (in the authoring container) — plain, keyword and multiple results, an op that
    raises, a result of the wrong type, ``-> None``, lazy and eager arguments, a lazy argument handed back as the
    result, a replica that is not the main process, unreadable inputs (eager: refused before the op runs and the
    main-pid marker stays, as there; lazy and untouched: never read).  Ours runs over a directory mount."""
    import importlib.util
    import json
    import logging
    from pathlib import Path

    from lzy_b200.api.v1 import startup
    from lzy_b200.serialization.registry import LzySerializerRegistry
    from lzy_b200.utils import tbpickle

    tbpickle.install()

    class DirMount:
        def __init__(self, directory):
            self.directory = directory

        def read(self, path, typ, serializers):
            with open(self.directory + path, "rb") as f:
                return serializers.find_serializer_by_type(typ).deserialize(f)

        def write(self, path, typ, data, serializers):
            with open(self.directory + path, "wb") as f:
                serializers.find_serializer_by_type(typ).serialize(data, f)

    def execute(mount, registry, op, args_paths, kwargs_paths, output_paths, exception_path, lazy, logger):
        startup.process_execution(registry, op, args_paths, kwargs_paths, output_paths, exception_path, logger, lazy,
                                  DirMount(mount))

    golden = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("startup_battery", golden / "startup_battery.py")
    battery = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(battery)
    logging.disable(logging.CRITICAL)
    try:
        ours = battery.run(dict(execute=execute), LzySerializerRegistry(), str(tmp_path))
    finally:
        logging.disable(logging.NOTSET)
    reference = json.loads((golden / "reference_startup_vectors.json").read_text())
    assert set(ours) == set(reference) and len(reference) == 13
    for tag, expected in reference.items():
        assert json.loads(json.dumps(ours[tag])) == expected, tag


def test_a_request_encoded_by_the_reference_dataclass_decodes_into_ours():
    """tests/golden/reference_request_vector.json: a ProcessingRequest built from pylzy's own dataclass and encoded by its
    own pickle() (in the authoring container) — what a stock client puts on a task's command line.  With
    install_as_lzy() the names in it (lzy.api.v1.startup.ProcessingRequest, lzy.serialization.registry.SerializerImport)
    resolve to this package's classes and every field arrives"""
    import json
    from pathlib import Path

    import lzy_b200
    from lzy_b200.api.v1.utils.pickle import unpickle
    from lzy_b200.serialization.registry import SerializerImport

    lzy_b200.install_as_lzy()
    vector = json.loads((Path(__file__).parent / "golden" / "reference_request_vector.json").read_text())
    req = unpickle(vector["text"])
    assert type(req) is ProcessingRequest and all(type(s) is SerializerImport for s in req.serializers)
    fields = vector["fields"]
    assert req.logger_config == fields["logger_config"] and req.op is len and req.lazy_arguments is fields["lazy_arguments"]
    assert [[s.module_name, s.class_name, s.priority] for s in req.serializers] == fields["serializers"]
    named = lambda typ, path: [typ.__name__, path]  # noqa: E731
    assert [named(*p) for p in req.args_paths] == fields["args_paths"]
    assert {k: named(*p) for k, p in req.kwargs_paths.items()} == fields["kwargs_paths"]
    assert [named(*p) for p in req.output_paths] == fields["output_paths"] and named(*req.exception_path) == fields["exception_path"]
