"""oracle/local_runtime.py — TEST INFRASTRUCTURE / TIMED CPU BASELINE ONLY.

CPU restatement of the reference's *local* @op data path, following it step by step
(paths relative to the reference tree):

  DefaultSnapshot.put_data     pylzy/lzy/api/v1/snapshot.py:141-172   serialize → HashingIO(MD5) tmpfile
                                                                         → uri = prefix + md5 → blob_exists → write
  HashingIO                    pylzy/lzy/api/v1/utils/hashing.py:10-38
  FsStorageClient.read/write   pylzy/lzy/storage/async_/fs.py:15-38    8096-byte loops
  DefaultSnapshot.get_data     pylzy/lzy/api/v1/snapshot.py:108-139   size → read to tmpfile → deserialize
  LocalRuntime.exec            pylzy/lzy/api/v1/local/runtime.py:42-85 DFS topological order, sequential
  LocalRuntime.__exec_call     pylzy/lzy/api/v1/local/runtime.py:87-179 tmpdir, storage→file per arg,
                                                                         Popen(python -u startup.py <b64 request>),
                                                                         file→storage per result, exception re-raise
  startup.process_execution    pylzy/lzy/api/v1/startup.py:109-170     (oracle/_startup.py)
  pickle/unpickle              pylzy/lzy/api/v1/utils/pickle.py:7-12   base64(cloudpickle)

Serializer: the reference delegates object↔bytes to the un-vendored ``serialzy`` (>=1.2,<2).  Its
bytes are PARITY-UNPINNED; this oracle uses cloudpickle for every type (pylzy itself uses
cloudpickle for the op request).  The subprocess here imports only cloudpickle/numpy, whereas the
reference's startup re-imports lzy + serialzy + logging config, so this baseline is, if anything,
*faster* than the real reference.
"""
from __future__ import annotations

import base64
import hashlib
import os
import subprocess
import sys
import tempfile
import uuid
from collections import defaultdict
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple
from urllib.parse import urlparse

import cloudpickle

from . import _tbpickle

_tbpickle.install()

FS_CHUNK = 8096  # fs.py:19,33
_STARTUP = str(Path(__file__).resolve().parent / "_startup.py")


def md5_of_str(uri: str) -> str:  # hashing.py:6-7
    return hashlib.md5(uri.encode("utf-8")).hexdigest()


class HashingIO:
    """hashing.py:10-38 — wraps a binary stream, MD5s everything written."""

    def __init__(self, io_stream):
        self.io_stream = io_stream
        self._md5 = hashlib.md5()

    @property
    def md5(self) -> str:
        return self._md5.hexdigest()

    def write(self, s: bytes) -> int:
        self._md5.update(s)
        return self.io_stream.write(s)

    def __getattr__(self, name):
        return getattr(self.__dict__["io_stream"], name)


def path_from_uri(uri: str) -> str:  # fs.py:53-56
    parsed = urlparse(uri)
    return os.path.abspath(os.path.join(parsed.netloc, parsed.path))


def _pump(source, sink) -> None:
    """the storage client's copy loop: 8096 bytes at a time until a read comes back empty (fs.py:19-25, 33-38)"""
    for piece in iter(lambda: source.read(FS_CHUNK), b""):
        sink.write(piece)


def fs_read(uri: str, dest) -> None:  # fs.py:15-25
    with open(path_from_uri(uri), "rb") as stored:
        _pump(stored, dest)


def fs_write(uri: str, data) -> None:  # fs.py:27-38
    target = Path(path_from_uri(uri))
    target.parent.absolute().mkdir(parents=True, exist_ok=True)
    with open(target, "wb") as stored:
        _pump(data, stored)


def pickle_request(obj: Any) -> str:  # utils/pickle.py:7-8
    return base64.b64encode(cloudpickle.dumps(obj)).decode("ascii")


@dataclass
class Entry:
    id: str
    name: str
    typ: Any
    storage_uri: Optional[str] = None
    data_hash: Optional[str] = None


@dataclass
class Call:
    """The slice of LzyCall the local runtime looks at (core/call.py:49-100)."""
    func: Callable
    arg_entry_ids: List[str]
    kwarg_entry_ids: Dict[str, str]
    entry_ids: List[str]
    exception_id: str
    lazy_arguments: bool = False
    id: str = field(default_factory=lambda: str(uuid.uuid4()))


class OracleSnapshot:
    """DefaultSnapshot over the FS storage client (snapshot.py:77-191)."""

    def __init__(self, storage_prefix: str):
        self.prefix = storage_prefix  # "<storage>/lzy_runs/<wf>/inputs/"
        self.entries: Dict[str, Entry] = {}

    def create_entry(self, name: str, typ: Any) -> Entry:
        e = Entry(str(uuid.uuid4()), name, typ)
        self.entries[e.id] = e
        return e

    def get(self, eid: str) -> Entry:
        return self.entries[eid]

    def put_data(self, eid: str, data: Any) -> None:  # snapshot.py:141-172
        entry = self.entries[eid]
        with tempfile.NamedTemporaryFile() as raw:
            f = HashingIO(raw)
            cloudpickle.dump(data, f)  # serializer.serialize(data, f)
            f.seek(0)
            entry.data_hash = f.md5
            entry.storage_uri = self.prefix + f.md5
            if not os.path.exists(path_from_uri(entry.storage_uri)):
                fs_write(entry.storage_uri, raw)

    def get_data(self, eid: str):  # snapshot.py:108-139; returns (found, value)
        entry = self.entries[eid]
        if entry.storage_uri is None or not os.path.exists(path_from_uri(entry.storage_uri)):
            return False, None
        with tempfile.NamedTemporaryFile() as f:
            fs_read(entry.storage_uri, f)
            f.seek(0)
            return True, cloudpickle.load(f)


class OracleLocalRuntime:
    """LocalRuntime (local/runtime.py) over OracleSnapshot."""

    def __init__(self, storage_uri: Optional[str] = None, workflow: str = "wf"):
        self._tmp = None
        if storage_uri is None:
            self._tmp = tempfile.TemporaryDirectory(prefix="lzy_oracle_storage_")
            storage_uri = "file://" + self._tmp.name
        self.storage_uri = storage_uri
        self.workflow = workflow
        self.snapshot = OracleSnapshot(f"{storage_uri}/lzy_runs/{workflow}/inputs/")
        self.spawned = 0

    def close(self):
        if self._tmp is not None:
            self._tmp.cleanup()

    # -- what LzyCall.__post_init__ + workflow._barrier do around the runtime --------------------
    def make_call(self, func: Callable, args: Sequence[Any], n_outputs: int = 1, kwargs: Optional[Dict[str, Any]] = None,
                  lazy_arguments: bool = False) -> Call:
        """args may be entry ids (str produced by a previous call: a proxy) or local values, which
        are put_data'd immediately (core/call.py:62-100)."""
        arg_ids: List[str] = []
        for i, a in enumerate(args):
            if isinstance(a, _Ref):
                arg_ids.append(a.eid)
            else:
                e = self.snapshot.create_entry(f"{func.__name__}.arg{i}", type(a))
                self.snapshot.put_data(e.id, a)
                arg_ids.append(e.id)
        kw_ids: Dict[str, str] = {}
        for k, a in (kwargs or {}).items():
            if isinstance(a, _Ref):
                kw_ids[k] = a.eid
            else:
                e = self.snapshot.create_entry(f"{func.__name__}.{k}", type(a))
                self.snapshot.put_data(e.id, a)
                kw_ids[k] = e.id
        outs = [self.snapshot.create_entry(f"{func.__name__}.return_{i}", object).id for i in range(n_outputs)]
        exc = self.snapshot.create_entry(f"{func.__name__}.exception", tuple).id
        call = Call(func, arg_ids, kw_ids, outs, exc, lazy_arguments)
        # workflow.__gen_results_uri_skip_cache (core/workflow.py:283-298)
        for i, eid in enumerate(outs):
            ent = self.snapshot.get(eid)
            ent.storage_uri = f"{self.storage_uri}/lzy_runs/{self.workflow}/ops/{func.__name__}/{call.id}/return_{i}"
            ent.data_hash = md5_of_str(ent.storage_uri)
        ent = self.snapshot.get(exc)
        ent.storage_uri = f"{self.storage_uri}/lzy_runs/{self.workflow}/ops/{func.__name__}/{call.id}/exception"
        ent.data_hash = md5_of_str(ent.storage_uri)
        return call

    def refs(self, call: Call) -> List["_Ref"]:
        return [_Ref(e) for e in call.entry_ids]

    # -- LocalRuntime.exec (local/runtime.py:42-85) ------------------------------------------------
    def exec(self, calls: List[Call]) -> List[str]:
        order = self.topo_order(calls)
        for c in order:
            self._exec_call(c)
        return [c.id for c in order]

    @staticmethod
    def topo_order(calls: List[Call]) -> List[Call]:
        """The order LocalRuntime.exec runs one barrier in (local/runtime.py:44-85): a depth-first walk over *entries*
        — from every result entry to the results of the calls that read it — taken in reverse finishing order, each
        call run at the first of its results.  Quirks kept: a walk's starting entry is not marked as reached (it may
        finish twice; the per-call de-duplication absorbs that), readers are visited positional arguments first.
        Pinned by tests/golden/reference_order_vectors.json (the reference's own exec on forty random graphs)."""
        readers: Dict[str, List[Call]] = defaultdict(list)
        produced_by: Dict[str, Call] = {}
        for call in calls:
            for eid in [*call.arg_entry_ids, *call.kwarg_entry_ids.values()]:
                readers[eid].append(call)
            for eid in call.entry_ids:
                produced_by[eid] = call
        reached: set = set()
        finished: List[str] = []

        def walk(eid: str) -> None:
            for reader in readers[eid]:
                for result in reader.entry_ids:
                    if result not in reached:
                        reached.add(result)
                        walk(result)
            finished.append(eid)

        for eid in list(produced_by):
            if eid not in reached:
                walk(eid)
        order: List[Call] = []
        scheduled: set = set()
        for eid in reversed(finished):
            call = produced_by[eid]
            if call.id not in scheduled:
                scheduled.add(call.id)
                order.append(call)
        return order

    def _exec_call(self, call: Call) -> None:  # local/runtime.py:87-179
        with tempfile.TemporaryDirectory("_lzy_call") as folder:
            arg_desc: List[Tuple[Any, str]] = []
            for eid in call.arg_entry_ids:
                entry = self.snapshot.get(eid)
                with open(folder + "/" + eid, "wb+") as f:
                    fs_read(entry.storage_uri, f)
                arg_desc.append((entry.typ, f"/{eid}"))
            kw_desc: Dict[str, Tuple[Any, str]] = {}
            for name, eid in call.kwarg_entry_ids.items():
                entry = self.snapshot.get(eid)
                with open(folder + "/" + eid, "wb+") as f:
                    fs_read(entry.storage_uri, f)
                kw_desc[name] = (entry.typ, f"/{eid}")
            ret_desc = []
            for eid in call.entry_ids:
                Path(folder + "/" + eid).touch()
                ret_desc.append((self.snapshot.get(eid).typ, f"/{eid}"))
            exc_eid = call.exception_id
            (Path(folder) / exc_eid).touch()
            request = dict(op=call.func, args_paths=arg_desc, kwargs_paths=kw_desc, output_paths=ret_desc,
                           exception_path=(tuple, f"/{exc_eid}"), lazy_arguments=call.lazy_arguments)
            env = os.environ.copy()
            env.pop("LZY_OP_MAIN_PID", None)
            env["LZY_MOUNT"] = folder
            main_module = sys.modules["__main__"]
            if hasattr(main_module, "__file__") and main_module.__file__:
                d = os.path.dirname(main_module.__file__)
                env["PYTHONPATH"] = f"{env['PYTHONPATH']}:{d}" if "PYTHONPATH" in env else d
            proc = subprocess.Popen([sys.executable, "-u", _STARTUP, pickle_request(request)], env=env,
                                    stdout=subprocess.PIPE)
            self.spawned += 1
            for line in iter(proc.stdout.readline, b""):
                sys.stdout.write(line.decode("utf-8"))
            proc.stdout.close()
            rc = proc.wait()
            if rc != 0:
                exc_entry = self.snapshot.get(exc_eid)
                with open(folder + "/" + exc_eid, "rb") as f:
                    fs_write(exc_entry.storage_uri, f)
                found, exception = self.snapshot.get_data(exc_eid)
                if found and exception is not None:
                    _typ, value, trace = exception
                    raise value.with_traceback(trace)
                raise RuntimeError(f"Error during execution of {call.func}")
            for eid in call.entry_ids:
                entry = self.snapshot.get(eid)
                with open(folder + "/" + eid, "rb") as f:
                    fs_write(entry.storage_uri, f)

    def materialize(self, ref: "_Ref") -> Any:
        found, value = self.snapshot.get_data(ref.eid)
        if not found:
            raise RuntimeError(f"Cannot materialize data with entry id {ref.eid}")
        return value


@dataclass(frozen=True)
class _Ref:
    """Stand-in for an lzy_proxy that has not been materialised: names a snapshot entry."""
    eid: str


def run_chain(funcs: Sequence[Callable], x: Any, runtime: Optional[OracleLocalRuntime] = None) -> Any:
    """y = f_n(...f_1(x)) through the oracle runtime: the reference's call stack §3.1."""
    own = runtime is None
    rt = runtime or OracleLocalRuntime()
    try:
        calls: List[Call] = []
        cur: Any = x
        for f in funcs:
            c = rt.make_call(f, [cur])
            calls.append(c)
            cur = rt.refs(c)[0]
        rt.exec(calls)
        return rt.materialize(cur)
    finally:
        if own:
            rt.close()
