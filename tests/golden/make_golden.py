#!/usr/bin/env python
"""Generates tests/golden/*.json.  Run ONLY in the authoring container, where /root/reference exists:

    python tests/golden/make_golden.py

* reference_leaf_vectors.json — produced by executing the reference's own pure-Python leaf modules
  (they import without the missing third parties): pylzy/lzy/api/v1/utils/hashing.py (HashingIO,
  md5_of_str) and pylzy/lzy/storage/async_/fs.py (FsStorageClient read/write 8096-byte loops,
  path_from_uri).  `lzy.api.v1` itself cannot be imported (serialzy, envzy, generated protobuf
  stubs are absent), so the modules are loaded by file path.
* reference_proxy_vectors.json — answers of the reference's own lazy proxy (pylzy/lzy/proxy/automagic.py, loaded
  by file path) to the probes of tests/golden/proxy_battery.py; tests/test_lazy_proxy.py holds ours to them.
* reference_materialize_vectors.json — how the reference's own lzy_proxy (api/v1/utils/proxy_adapter.py) obtains a value over a
  fake workflow (tests/golden/materialize_battery.py): supplied value, snapshot, one barrier, the raised reasons.
* reference_types_vectors.json — answers of the reference's own annotation helpers (pylzy/lzy/api/v1/utils/types.py) to the
  questions of tests/golden/types_battery.py; tests/test_call_signatures.py holds ours to them.
* reference_call_vectors.json — what the reference's own argument binding (pylzy/lzy/core/call.py:
  infer_and_validate_call_signature) does with the calls of tests/golden/call_battery.py: refusals with their messages,
  entry names, types, local data.
* reference_lzycall_vectors.json — what constructing the reference's own LzyCall does to a recording snapshot for the calls of
  tests/golden/lzycall_battery.py: entries created, local arguments written at construction, proxies reusing their
  producers' entries, the environment combined Lzy → workflow → call.
* reference_snapshot_vectors.json — the reference's own DefaultSnapshot on its own FsStorageClient over a temporary directory, for
  the script of tests/golden/snapshot_battery.py: content addresses (prefix + md5 of the serialized bytes), equal data
  stored once, what reading answers; the oracle's restatement (oracle/local_runtime.py) is held to it.
* reference_order_vectors.json — the order in which the reference's own LocalRuntime.exec runs the calls of forty seeded random
  graphs (tests/golden/order_battery.py); the oracle's topo_order is held to it.
* reference_file_vectors.json — the reference's own FileSerializer (pylzy/lzy/serialization/file.py) on the files and streams of
  tests/golden/file_battery.py: the raw_file byte format and how streams with / without / shorter than the header read.
* reference_wrapper_vectors.json — the reference's own @op decorator and LazyCallWrapper (core/op.py, core/call.py) on
  tests/golden/wrapper_battery.py: what the decorator records, the unannotated refusal, plain calls outside a workflow,
  proxies / None / eager values inside one.
* reference_lifecycle_vectors.json — the reference's own LzyWorkflow entered, used and left over a recording runtime
  (tests/golden/lifecycle_battery.py): start / exec(n) / finish / abort in order, what is raised, the active workflow,
  the snapshot prefix.
* reference_uri_vectors.json — the storage URIs (and entry hashes) the reference's own LzyWorkflow assigns to the results and
  exceptions of the calls of tests/golden/uri_battery.py, without and with LZY_USER: the cache naming contract.
* reference_startup_vectors.json — the reference's own op-side entry point (pylzy/lzy/api/v1/startup.py: process_execution)
  run over slot files for the scenarios of tests/golden/startup_battery.py: which slots get written, what the
  exception slot holds, what is raised, what is left of LZY_OP_MAIN_PID.
* reference_request_vector.json — a ProcessingRequest built from the reference's own dataclass and encoded by its own pickle():
  the text a stock client puts on a task's command line; after install_as_lzy() ours must decode it into its own
  ProcessingRequest with the same fields.
* reference_registry_vectors.json — the reference's own LzySerializerRegistry subclass (serialization/registry.py) over a stand-in base
  on tests/golden/registry_battery.py: user serializers shipped as imports, load_imports, the __main__ ban, copyreg.
* reference_pickle_vectors.json — what the reference's request encoding (pylzy/lzy/api/v1/utils/pickle.py) emits
  for a few objects; ours must decode them and encode the same text.
* digest_vectors.json — known answers of the data plane's digest from the oracle's C restatement
  cross-checked with the pure-Python definition in tests/test_abi_and_oracle.py.
"""
import asyncio
import hashlib
import importlib.util
import io
import json
import os
import sys
import tempfile
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/pylzy")


def load(path: Path, name: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    import oracle
    out = {}
    hashing = load(REF / "lzy/api/v1/utils/hashing.py", "ref_hashing")
    out["md5_of_str"] = [{"in": s, "out": hashing.md5_of_str(s)} for s in
                         ["", "file:///tmp/lzy_storage/lzy_runs/wf/ops/foo/1234/return_0", "ʎzy"]]
    hio = []
    for writes in ([b""], [b"Hello"], [b"a" * 8096, b"b" * 8096, b"tail"], [bytes(range(256)) * 33]):
        sink = io.BytesIO()
        h = hashing.HashingIO(sink)
        for w in writes:
            h.write(w)
        assert sink.getvalue() == b"".join(writes)
        hio.append({"writes_hex": [w.hex() for w in writes], "md5": h.md5})
    out["hashing_io"] = hio

    # fs.py imports lzy.storage.api (pure) — make the package importable without lzy/__init__ side effects
    sys.path.insert(0, str(REF))
    fs = load(REF / "lzy/storage/async_/fs.py", "ref_fs")
    client = fs.FsStorageClient()
    fsv = []
    for nbytes, seed in ((0, 1), (1, 2), (8095, 3), (8096, 4), (8097, 5), (1 << 20, 42), (3 * 8096 + 17, 6)):
        data = oracle.pattern(nbytes, seed).tobytes()
        with tempfile.TemporaryDirectory() as d:
            uri = "file://" + os.path.join(d, "x", "blob")
            calls = []
            asyncio.run(client.write(uri, io.BytesIO(data), progress=lambda n, r: calls.append(n)))
            dst = io.BytesIO()
            asyncio.run(client.read(uri, dst))
            assert asyncio.run(client.size_in_bytes(uri)) == nbytes
            assert asyncio.run(client.blob_exists(uri))
        fsv.append({"nbytes": nbytes, "seed": seed, "src_md5": hashlib.md5(data).hexdigest(),
                    "dst_md5": hashlib.md5(dst.getvalue()).hexdigest(), "loop_iterations": len(calls)})
    out["fs_copy"] = fsv
    p = fs.FsStorageClient._FsStorageClient__path_from_uri
    out["path_from_uri"] = [{"in": u, "out": p(u)} for u in
                            ["file:///tmp/lzy_storage/a", "file:///tmp/lzy_storage/lzy_runs/wf/inputs/abc"]]
    (ROOT / "tests/golden/reference_leaf_vectors.json").write_text(json.dumps(out, indent=1))

    proxy_battery = load(ROOT / "tests/golden/proxy_battery.py", "proxy_battery")  # (REF/tests shadows ours on sys.path)
    automagic = load(REF / "lzy/proxy/automagic.py", "ref_automagic")
    answers = json.loads(json.dumps(proxy_battery.run(automagic.proxy), default=str))
    (ROOT / "tests/golden/reference_proxy_vectors.json").write_text(json.dumps(answers, indent=1, sort_keys=True))

    # the annotation helpers: the reference's own module, its absent third party `serialzy` replaced by two-name stubs
    # (the helpers only name SerializerRegistry in a signature and compare against EmptyContent)
    serialzy_stub, api_stub, types_stub = types.ModuleType("serialzy"), types.ModuleType("serialzy.api"), types.ModuleType("serialzy.types")
    api_stub.SerializerRegistry = type("SerializerRegistry", (), {})
    types_stub.EmptyContent = type("EmptyContent", (), {})
    sys.modules.update({"serialzy": serialzy_stub, "serialzy.api": api_stub, "serialzy.types": types_stub})
    ref_types = load(REF / "lzy/api/v1/utils/types.py", "ref_types")
    ref_result = load(REF / "lzy/proxy/result.py", "ref_result")
    from lzy.proxy.result import Absence as RefAbsence, Result as RefResult  # the classes ref_types itself imported
    types_battery = load(ROOT / "tests/golden/types_battery.py", "types_battery")
    answers = types_battery.run(dict(
        infer_real_types=ref_types.infer_real_types, infer_return_type=ref_types.infer_return_type,
        get_default_args=ref_types.get_default_args, is_subtype=ref_types.is_subtype,
        check_types_serialization_compatible=ref_types.check_types_serialization_compatible,
        EmptyContent=types_stub.EmptyContent, Absence=RefAbsence, Result=RefResult))
    del ref_result
    (ROOT / "tests/golden/reference_types_vectors.json").write_text(json.dumps(answers, indent=1, sort_keys=True))

    # argument binding: the reference's own lzy/core/call.py.  Its imports of the environment / logging / snapshot
    # layers are stubbed (they only supply names for annotations and base classes); serialzy's get_type and
    # EmptyContent come from this package, the same ones our side of the battery uses
    import logging
    from lzy_b200.serialization import types as our_types
    from lzy_b200.serialization.registry import LzySerializerRegistry

    def stub(name, **attrs):
        module = types.ModuleType(name)
        module.__dict__.update(attrs)
        sys.modules[name] = module
        return module

    def load_as(path, name):
        spec = importlib.util.spec_from_file_location(name, path)
        module = importlib.util.module_from_spec(spec)
        sys.modules[name] = module
        spec.loader.exec_module(module)
        return module

    types_stub.EmptyContent, types_stub.get_type = our_types.EmptyContent, our_types.get_type
    api_stub.Schema = type("Schema", (), {})
    for name, attrs in {
        "lzy.env": {}, "lzy.env.provisioning": {}, "lzy.env.provisioning.provisioning": {"Provisioning": object},
        "lzy.env.python": {}, "lzy.env.python.base": {"ModulePathsList": list, "PackagesDict": dict},
        "lzy.env.container": {}, "lzy.env.container.base": {"BaseContainer": object},
        "lzy.env.environment": {"LzyEnvironment": type("LzyEnvironment", (), {}), "EnvVarsType": dict},
        "lzy.env.mixin": {"WithEnvironmentMixin": type("WithEnvironmentMixin", (), {})},
        "lzy.logs": {}, "lzy.logs.config": {"get_logger": logging.getLogger},
        "lzy.api": {}, "lzy.api.v1": {}, "lzy.api.v1.snapshot": {"Snapshot": object}, "lzy.api.v1.utils": {},
    }.items():
        stub(name, **attrs)
    load_as(REF / "lzy/api/v1/signatures.py", "lzy.api.v1.signatures")
    load_as(REF / "lzy/api/v1/utils/types.py", "lzy.api.v1.utils.types")
    ref_adapter = load_as(REF / "lzy/api/v1/utils/proxy_adapter.py", "lzy.api.v1.utils.proxy_adapter")
    ref_call = load_as(REF / "lzy/core/call.py", "ref_call")
    call_battery = load(ROOT / "tests/golden/call_battery.py", "call_battery")
    bound = call_battery.run(dict(bind=ref_call.infer_and_validate_call_signature, lzy_proxy=ref_adapter.lzy_proxy,
                                  materialize=ref_adapter.materialize, is_lzy_proxy=ref_adapter.is_lzy_proxy,
                                  Result=RefResult, registry=LzySerializerRegistry()))
    (ROOT / "tests/golden/reference_call_vectors.json").write_text(json.dumps(bound, indent=1, sort_keys=True))

    # touching a proxy: the reference's own lzy_proxy over a fake workflow (snapshot answers scripted, barrier recorded)
    from lzy.proxy.result import Absence as RefAbsence2
    materialize_battery = load(ROOT / "tests/golden/materialize_battery.py", "materialize_battery")
    touched = materialize_battery.run(dict(
        lzy_proxy=ref_adapter.lzy_proxy, materialize=ref_adapter.materialize, materialized=ref_adapter.materialized,
        is_lzy_proxy=ref_adapter.is_lzy_proxy, get_proxy_entry_id=ref_adapter.get_proxy_entry_id, Result=RefResult,
        Absence=RefAbsence2))
    (ROOT / "tests/golden/reference_materialize_vectors.json").write_text(json.dumps(touched, indent=1, sort_keys=True))

    # recording a call: the reference's own LzyCall over a recording snapshot
    lzycall_battery = load(ROOT / "tests/golden/lzycall_battery.py", "lzycall_battery")
    recorded = lzycall_battery.run(dict(bind=ref_call.infer_and_validate_call_signature, LzyCall=ref_call.LzyCall,
                                        lzy_proxy=ref_adapter.lzy_proxy, Result=RefResult, registry=LzySerializerRegistry()))
    (ROOT / "tests/golden/reference_lzycall_vectors.json").write_text(json.dumps(recorded, indent=1, sort_keys=True))

    # @op and calling the wrapper: the reference's own decorator and LazyCallWrapper; the active workflow is a fake
    class BatteryEnv:
        def combine(self, other):
            return self

        def validate(self):
            pass

    class ActiveWorkflow:
        instance = None

        @classmethod
        def get_active(cls):
            return cls.instance

    sys.modules["lzy.env.environment"].LzyEnvironment = BatteryEnv
    stub("lzy.core")
    ref_call_module = load_as(REF / "lzy/core/call.py", "lzy.core.call")
    stub("lzy.core.workflow", LzyWorkflow=ActiveWorkflow)
    stub("lzy.utils.format", pretty_function=load(REF / "lzy/utils/format.py", "ref_format").pretty_function)
    ref_op = load_as(REF / "lzy/core/op.py", "ref_op")
    wrapper_battery = load(ROOT / "tests/golden/wrapper_battery.py", "wrapper_battery")
    wrapped = wrapper_battery.run(dict(op=ref_op.op, LzyWorkflow=ActiveWorkflow, is_lzy_proxy=ref_adapter.is_lzy_proxy,
                                       get_proxy_entry_id=ref_adapter.get_proxy_entry_id, Result=RefResult, Env=BatteryEnv,
                                       registry=LzySerializerRegistry()))
    del ref_call_module
    del sys.modules["lzy.core.workflow"]
    (ROOT / "tests/golden/reference_wrapper_vectors.json").write_text(json.dumps(wrapped, indent=1, sort_keys=True))

    # result URIs: the reference's own LzyWorkflow (whiteboard / logging / environment imports stubbed), its two
    # private URI rules driven directly over the battery's fake calls and snapshot
    from types import SimpleNamespace

    class WithLogger:
        def __post_init__(self):
            pass

        @property
        def log(self):
            return logging.getLogger("reference")

    for name, attrs in {
        "ai": {}, "ai.lzy": {}, "ai.lzy.v1": {}, "ai.lzy.v1.whiteboard": {}, "ai.lzy.v1.whiteboard.whiteboard_pb2": {"Whiteboard": object},
        "lzy.api.v1.remote": {}, "lzy.api.v1.remote.lzy_service_client": {"USER_ENV": "LZY_USER"},
        "lzy.api.v1.whiteboards": {"WritableWhiteboard": object}, "lzy.env.base": {"WithLogger": WithLogger},
    }.items():
        stub(name, **attrs)
    sys.modules["lzy.api.v1.snapshot"].DefaultSnapshot = object
    load_as(REF / "lzy/api/v1/utils/hashing.py", "lzy.api.v1.utils.hashing")
    load_as(REF / "lzy/api/v1/utils/validation.py", "lzy.api.v1.utils.validation")
    ref_workflow = load_as(REF / "lzy/core/workflow.py", "ref_workflow")
    uri_battery = load(ROOT / "tests/golden/uri_battery.py", "uri_battery")

    def ref_make(name, storage, snapshot):
        w = ref_workflow.LzyWorkflow(name=name, owner=SimpleNamespace(storage_uri=storage), env=None)
        w._LzyWorkflow__snapshot = snapshot
        return w

    def ref_assign(w, call):
        (w._LzyWorkflow__gen_results_uri_in_cache if call.cache else w._LzyWorkflow__gen_results_uri_skip_cache)(call)

    # entering / using / leaving a workflow: the same reference class driving a recording runtime
    lifecycle_battery = load(ROOT / "tests/golden/lifecycle_battery.py", "lifecycle_battery")

    class RecordingSnapshot:
        def __init__(self, registry, prefix, client, storage_name):
            self.prefix, self.entries = prefix, {}

        def get(self, eid):
            return self.entries[eid]

    ref_workflow.DefaultSnapshot = RecordingSnapshot

    def ref_make_started(name, runtime, log, eager):
        owner = SimpleNamespace(storage_uri="s3://bucket/prefix", runtime=runtime, serializer_registry=None,
                                storage_client=None, storage_name="storage")
        return ref_workflow.LzyWorkflow(name=name, owner=owner, env=None, eager=eager)

    lifecycles = []
    for user in (None, "igor"):
        os.environ.pop("LZY_USER", None)
        if user is not None:
            os.environ["LZY_USER"] = user
        lifecycles.append(lifecycle_battery.run(dict(
            make_workflow=ref_make_started, active=lambda: getattr(ref_workflow.LzyWorkflow.get_active(), "name", None)), user))
    os.environ.pop("LZY_USER", None)
    (ROOT / "tests/golden/reference_lifecycle_vectors.json").write_text(json.dumps(lifecycles, indent=1, sort_keys=True))

    uris = []
    for user in (None, "igor"):
        os.environ.pop("LZY_USER", None)
        if user is not None:
            os.environ["LZY_USER"] = user
        uris.append(uri_battery.run(dict(make_workflow=ref_make, assign=ref_assign), user))
    os.environ.pop("LZY_USER", None)
    (ROOT / "tests/golden/reference_uri_vectors.json").write_text(json.dumps(uris, indent=1, sort_keys=True))

    # the op-side entry point: the reference's own process_execution over slot files in a temporary LZY_MOUNT, with
    # this package's serializer registry on both sides (serialzy is absent)
    from lzy_b200.serialization.registry import SerializerImport
    from lzy_b200.utils import tbpickle
    tbpickle.install()  # sys.exc_info() goes into the exception slot, traceback included
    load_as(REF / "lzy/api/v1/utils/pickle.py", "lzy.api.v1.utils.pickle")
    sys.modules["lzy.logs.config"].configure_logging = lambda config: None
    sys.modules["lzy.logs.config"].get_remote_logger = logging.getLogger
    stub("lzy.serialization")
    stub("lzy.serialization.registry", LzySerializerRegistry=LzySerializerRegistry, SerializerImport=SerializerImport)
    ref_startup = load_as(REF / "lzy/api/v1/startup.py", "ref_startup")
    startup_battery = load(ROOT / "tests/golden/startup_battery.py", "startup_battery")

    def ref_execute(mount, registry, op, args_paths, kwargs_paths, output_paths, exception_path, lazy, logger):
        os.environ["LZY_MOUNT"] = mount
        try:
            ref_startup.process_execution(registry, op, args_paths, kwargs_paths, output_paths, exception_path, logger, lazy)
        finally:
            os.environ.pop("LZY_MOUNT", None)

    logging.disable(logging.CRITICAL)
    with tempfile.TemporaryDirectory() as d:
        executed = startup_battery.run(dict(execute=ref_execute), LzySerializerRegistry(), d)
    logging.disable(logging.NOTSET)
    (ROOT / "tests/golden/reference_startup_vectors.json").write_text(json.dumps(executed, indent=1, sort_keys=True))

    # the workflow's data table: the reference's own DefaultSnapshot over its own FsStorageClient in a temporary directory
    import contextlib
    sys.modules["lzy.logs.config"].get_syslog_color = lambda: "white"
    ref_snapshot = load_as(REF / "lzy/api/v1/snapshot.py", "ref_snapshot")
    snapshot_battery = load(ROOT / "tests/golden/snapshot_battery.py", "snapshot_battery")
    with tempfile.TemporaryDirectory() as d:
        prefix, inputs = "file://" + d + "/lzy_runs/wf/inputs/", d + "/lzy_runs/wf/inputs"
        snap = ref_snapshot.DefaultSnapshot(LzySerializerRegistry(), prefix, fs.FsStorageClient(), "storage")

        def quiet(coro):  # the progress bars go to stdout
            with contextlib.redirect_stdout(io.StringIO()):
                return asyncio.run(coro)

        def get(entry):
            answer = quiet(snap.get_data(entry.id))
            return ("result", repr(answer.value)) if isinstance(answer, RefResult) else ("absence", str(answer.cause))

        def attr(entry, name):
            try:
                return getattr(entry, name)
            except ValueError:
                return None

        table = snapshot_battery.run(dict(
            create=snap.create_entry, put=lambda entry, value: quiet(snap.put_data(entry.id, value)), get=get,
            uri=lambda entry: attr(entry, "storage_uri"), hash=lambda entry: attr(entry, "data_hash"),
            set_uri=lambda entry, uri: setattr(entry, "storage_uri", uri),
            stored=lambda: sorted(os.listdir(inputs)) if os.path.isdir(inputs) else []), prefix)
    (ROOT / "tests/golden/reference_snapshot_vectors.json").write_text(json.dumps(table, indent=1, sort_keys=True))

    # execution order of one barrier: the reference's own LocalRuntime.exec, its per-call execution replaced by a recorder
    stub("lzy.api.v1.local")
    load_as(REF / "lzy/api/v1/exceptions.py", "lzy.api.v1.exceptions")
    load_as(REF / "lzy/api/v1/runtime.py", "lzy.api.v1.runtime")
    sys.modules["lzy.api.v1.startup"] = ref_startup
    sys.modules["lzy.logs.config"].get_logging_config = dict
    sys.modules["lzy.logs.config"].COLOURS, sys.modules["lzy.logs.config"].RESET_COLOR = {}, ""
    ref_local = load_as(REF / "lzy/api/v1/local/runtime.py", "ref_local_runtime")
    order_battery = load(ROOT / "tests/golden/order_battery.py", "order_battery")

    def ref_order(calls):
        runtime = ref_local.LocalRuntime()
        runtime._LocalRuntime__workflow = object()
        seen = []

        async def record(call):
            seen.append(call.id)

        runtime._LocalRuntime__exec_call = record
        asyncio.run(runtime.exec(calls, lambda step: None))
        return seen

    (ROOT / "tests/golden/reference_order_vectors.json").write_text(
        json.dumps(order_battery.run(dict(order=ref_order)), indent=1, sort_keys=True))

    # File payloads: the reference's own FileSerializer (its serialzy base class stubbed: only _check_types_valid is used)
    import collections
    import enum
    api_stub.Schema = collections.namedtuple("Schema", "data_format schema_format meta")
    api_stub.StandardSchemaFormats = enum.Enum("StandardSchemaFormats", "no_schema")
    api_stub.VersionBoundary = object
    stub("serialzy.base", DefaultSchemaSerializerByReference=type("DefaultSchemaSerializerByReference", (), {
        "_check_types_valid": lambda self, schema_type, user_type: None}))
    for name in ("ai.lzy.v1.workflow",):
        stub(name)
    stub("ai.lzy.v1.workflow.workflow_pb2", VmPoolSpec=object)
    ref_lzy_types = load_as(REF / "lzy/types.py", "lzy.types")
    ref_file = load_as(REF / "lzy/serialization/file.py", "ref_file_serializer")
    file_battery = load(ROOT / "tests/golden/file_battery.py", "file_battery")
    ref_file_serializer = ref_file.FileSerializer()
    with tempfile.TemporaryDirectory() as d:
        files = file_battery.run(dict(File=ref_lzy_types.File, serialize=ref_file_serializer._serialize,
                                      deserialize=lambda source: ref_file_serializer._deserialize(source, ref_lzy_types.File)), d)
    (ROOT / "tests/golden/reference_file_vectors.json").write_text(json.dumps(files, indent=1, sort_keys=True))

    # a task's command-line request as a stock client pickles it: the reference's own ProcessingRequest dataclass, its module
    # registered under its real name so that the pickle refers to lzy.api.v1.startup, encoded by the reference's own
    # pickle().  (SerializerImport's home, lzy/serialization/registry.py, cannot be loaded without serialzy: a dataclass
    # with the same three fields and the same module path stands in — a pickle holds nothing else of it.)
    import dataclasses

    @dataclasses.dataclass
    class SerializerImportStandIn:
        module_name: str
        class_name: str
        priority: int

    SerializerImportStandIn.__module__, SerializerImportStandIn.__qualname__ = "lzy.serialization.registry", "SerializerImport"
    SerializerImportStandIn.__name__ = "SerializerImport"
    sys.modules["lzy.serialization.registry"].SerializerImport = SerializerImportStandIn
    stock_startup = load_as(REF / "lzy/api/v1/startup.py", "lzy.api.v1.startup")
    stock_request = stock_startup.ProcessingRequest(
        {"version": 1}, [SerializerImportStandIn("some.module", "SomeSerializer", 7)], len, [(list, "/call-1/e-1")],
        {"k": (int, "/call-1/e-2")}, [(int, "/call-1/e-3")], (tuple, "/call-1/e-4"), lazy_arguments=False)
    (ROOT / "tests/golden/reference_request_vector.json").write_text(json.dumps({
        "text": sys.modules["lzy.api.v1.utils.pickle"].pickle(stock_request),
        "fields": {"logger_config": {"version": 1}, "serializers": [["some.module", "SomeSerializer", 7]], "op": "builtins.len",
                   "args_paths": [["list", "/call-1/e-1"]], "kwargs_paths": {"k": ["int", "/call-1/e-2"]},
                   "output_paths": [["int", "/call-1/e-3"]], "exception_path": ["tuple", "/call-1/e-4"], "lazy_arguments": False}},
        indent=1))

    # lzy.serializer_registry: the reference's own subclass over a stand-in for serialzy's DefaultSerializerRegistry (this
    # package's base registry plus the empty _type_registry the subclass iterates) — what is pinned is the subclass's own
    # logic: user serializers tracked and shipped as imports, the __main__ ban, copyreg hooks
    from lzy_b200.serialization.api import SerializerRegistry as OurBaseRegistry

    class StandInDefaultRegistry(OurBaseRegistry):
        def __init__(self):
            super().__init__()
            self._type_registry = {}

    api_stub.Serializer = object
    stub("serialzy.registry", DefaultSerializerRegistry=StandInDefaultRegistry)
    sys.modules["serialzy.base"].DefaultSchemaSerializerByReference = type("DefaultSchemaSerializerByReference", (), {
        "_check_types_valid": lambda self, schema_type, user_type: None,
        "handles": lambda self, typ: typ == self.supported_types(),
        "serialize": lambda self, obj, dest: self._serialize(obj, dest),
        "deserialize": lambda self, source, typ=None: self._deserialize(source, typ)})
    load_as(REF / "lzy/serialization/file.py", "lzy.serialization.file")
    ref_registry = load_as(REF / "lzy/serialization/registry.py", "ref_registry")
    registry_battery = load(ROOT / "tests/golden/registry_battery.py", "registry_battery")
    (ROOT / "tests/golden/reference_registry_vectors.json").write_text(
        json.dumps(registry_battery.run(dict(new_registry=ref_registry.LzySerializerRegistry)), indent=1, sort_keys=True))

    ref_pickle = load(REF / "lzy/api/v1/utils/pickle.py", "ref_pickle")
    objs = [None, 42, -1.5, "ʎzy", b"\x00\xff", [1, "two", 3.0], {"k": (1, 2), "n": None}, (str, "/call/entry"), {1, 2, 3}]
    pv = []
    for o in objs:
        text = ref_pickle.pickle(o)
        assert ref_pickle.unpickle(text) == o
        pv.append({"repr": repr(o), "text": text})
    (ROOT / "tests/golden/reference_pickle_vectors.json").write_text(json.dumps(pv, indent=1))

    vecs = [{"name": "empty", "hex": ""}, {"name": "Hello", "hex": b"Hello".hex()},
            {"name": "one-word", "hex": bytes(range(8)).hex()}, {"name": "17-bytes", "hex": bytes(range(17)).hex()},
            {"name": "zeros-4096", "hex": (b"\0" * 4096).hex()}]
    for v in vecs:
        v["digest"] = oracle.digest_hex(bytes.fromhex(v["hex"]))
    for nbytes, seed in ((1 << 20, 7), ((2 << 20) + 5, 1234), (64 << 20, 7)):
        vecs.append({"name": f"pattern-{nbytes}-{seed}", "pattern_nbytes": nbytes, "pattern_seed": seed,
                     "digest": oracle.digest_hex(oracle.pattern(nbytes, seed))})
    (ROOT / "tests/golden/digest_vectors.json").write_text(json.dumps({"vectors": vecs}, indent=1))
    print("golden vectors written")


if __name__ == "__main__":
    main()
