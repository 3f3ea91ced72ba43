"""`FrameSnapshot` — `DefaultSnapshot`'s contract (pylzy/lzy/api/v1/snapshot.py:93-191) as the reference's
pylzy/tests/api/v1/test_snapshot.py:38-86 pins it, plus what is new here: HBM is not a disk, so an entry holds its
frame only while a queued call or a live proxy can still ask for it (pin / unpin / produced / persist).  The
store is a dict-backed stand-in with the `FrameStore` surface the snapshot uses; the real one is exercised on the
GPU by tests/test_workflow_gpu.py."""
import asyncio
import hashlib
import pickle
import uuid

import pytest

from lzy_b200.api.v1.snapshot import FrameSnapshot
from lzy_b200.proxy_result import Absence, Result
from lzy_b200.serialization.api import Serializer, SerializerRegistry
from lzy_b200.serialization.registry import LzySerializerRegistry


class DictStore:
    def __init__(self):
        self.data, self.refs, self.durable, self.budget_checks = {}, {}, set(), 0

    def put(self, prefix, data, typ, registry, name=""):
        raw = pickle.dumps(data)
        h = hashlib.md5(raw).hexdigest()
        self.data.setdefault(prefix + h, data)
        return h

    def exists(self, uri):
        return uri in self.data

    def get(self, uri, typ, registry):
        return self.data[uri]

    def alias(self, src, dst):
        self.data[dst] = self.data[src]

    def retain(self, uri):
        self.refs[uri] = self.refs.get(uri, 0) + 1

    def release(self, uri):
        self.refs[uri] -= 1
        if self.refs[uri] == 0:
            del self.refs[uri]
            self.data.pop(uri, None)

    def persist(self, uri):
        self.durable.add(uri)

    def enforce_budget(self, protect=None):
        self.budget_checks += 1


@pytest.fixture
def snap():
    store = DictStore()
    return FrameSnapshot(LzySerializerRegistry(), "hbm://bucket/lzy_runs/wf/inputs/", store), store


def run(coro):
    return asyncio.run(coro)


def test_put_get_and_content_addressing(snap):
    s, store = snap
    e1, e2 = s.create_entry("name", str), s.create_entry("other", str)
    run(s.put_data(e1.id, "some_str"))
    run(s.put_data(e2.id, "some_str"))
    assert run(s.get_data(e1.id)) == Result("some_str")
    assert e1.storage_uri == e2.storage_uri == "hbm://bucket/lzy_runs/wf/inputs/" + e1.data_hash
    assert len(store.data) == 1 and store.refs[e1.storage_uri] == 2  # equal data: one blob, two claims


def test_unknown_entries_and_not_uploaded_data(snap):
    s, _ = snap
    with pytest.raises(ValueError, match="does not exist"):
        run(s.get_data(str(uuid.uuid4())))
    with pytest.raises(ValueError, match="does not exist"):
        run(s.put_data(str(uuid.uuid4()), "some_str"))
    e = s.create_entry("name", str)
    assert isinstance(run(s.get_data(e.id)), Absence)  # no URI yet
    e.storage_uri = "hbm://bucket/name"
    assert isinstance(run(s.get_data(e.id)), Absence)  # a URI nobody wrote to
    with pytest.raises(ValueError, match="is not loaded to storage"):
        run(s.copy_data(e.id, "hbm://bucket/elsewhere"))


def test_put_overrides_a_preassigned_uri(snap):
    s, _ = snap
    e = s.create_entry("name", str)
    e.storage_uri = "hbm://bucket/name"
    e.storage_uri = "hbm://bucket/name2"
    run(s.put_data(e.id, "some_str"))
    assert run(s.get_data(e.id)) == Result("some_str")


def test_copy_data_aliases_the_blob(snap):
    s, store = snap
    e = s.create_entry("name", int)
    run(s.put_data(e.id, 42))
    run(s.copy_data(e.id, "hbm://bucket/whiteboard/field"))
    assert store.data["hbm://bucket/whiteboard/field"] == 42


class _NotAvailable(Serializer):
    def supported_types(self):
        return str

    def available(self):
        return False

    def data_format(self):
        return "never"

    def _serialize(self, obj, dest):
        raise AssertionError

    def _deserialize(self, source, schema_type, user_type=None):
        raise AssertionError


def test_entries_need_a_usable_serializer():
    with pytest.raises(TypeError, match="Cannot find serializer for type"):
        FrameSnapshot(SerializerRegistry(), "hbm://x/", DictStore()).create_entry("name", str)
    reg = SerializerRegistry()
    reg.register_serializer(_NotAvailable())
    with pytest.raises(TypeError, match="is not available, please install"):
        FrameSnapshot(reg, "hbm://x/", DictStore()).create_entry("name", str)


# ---- lifetime: an entry's claim on its frame -------------------------------------------------------------
def test_local_data_is_released_when_the_last_pin_goes(snap):
    s, store = snap
    e = s.create_entry("arg", str)
    s.pin(e.id)            # a queued call consumes it
    s.pin(e.id)            # … and so does another
    run(s.put_data(e.id, "payload"))
    uri = e.storage_uri
    s.unpin(e.id)
    assert store.exists(uri)
    s.unpin(e.id)
    assert not store.exists(uri) and uri not in store.refs


def test_op_result_waits_for_its_readers_then_goes(snap):
    s, store = snap
    e = s.create_entry("ret", str)
    e.storage_uri = "hbm://bucket/lzy_runs/wf/ops/f/call-1/return_0"
    s.pin(e.id)                                  # the caller's lazy proxy
    store.data[e.storage_uri] = "result"         # the op ran: the slots runtime registered the frame
    s.produced(e.id)
    assert store.refs[e.storage_uri] == 1 and store.budget_checks == 1
    assert run(s.get_data(e.id)) == Result("result")
    s.unpin(e.id)                                # the proxy died
    assert not store.exists(e.storage_uri)


def test_result_nobody_waits_for_is_dropped_as_soon_as_it_exists(snap):
    s, store = snap
    e = s.create_entry("ret", str)
    e.storage_uri = "hbm://bucket/r"
    store.data[e.storage_uri] = "result"
    s.produced(e.id)
    assert not store.exists(e.storage_uri)


def test_cached_results_persist_and_go_to_the_durable_tier(snap):
    s, store = snap
    e = s.create_entry("ret", str)
    e.storage_uri = "hbm://bucket/lzy_runs/wf/ops/f/1.0/abc/return_0"
    s.persist(e.id)
    store.data[e.storage_uri] = "result"
    s.produced(e.id)
    assert store.exists(e.storage_uri) and e.storage_uri in store.durable
    s.pin(e.id)
    s.unpin(e.id)
    assert store.exists(e.storage_uri)  # kept until the runtime is cleared


def test_retain_all_keeps_everything(snap):
    s, store = snap
    s.auto_release = False
    e = s.create_entry("arg", str)
    s.pin(e.id)
    run(s.put_data(e.id, "payload"))
    s.unpin(e.id)
    assert store.exists(e.storage_uri)
