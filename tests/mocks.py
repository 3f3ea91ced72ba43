"""Test doubles for the Python API, in the spirit of pylzy/tests/api/v1/mocks.py:70-99 (RuntimeMock,
StorageClientMock): a snapshot that keeps cloudpickled bytes in a dict and a runtime that records the
calls it was asked to execute (and, optionally, executes them in process, sequentially, in the local
runtime's DFS-topological order).  TEST INFRASTRUCTURE — the product has no such path."""
from __future__ import annotations

import hashlib
import sys
import uuid
from typing import Any, Callable, Dict, List, Optional, Type

import cloudpickle

from lzy_b200.api.v1.runtime import ProgressStep, Runtime
from lzy_b200.api.v1.snapshot import Snapshot, SnapshotEntry
from lzy_b200.proxy_result import Absence, Result
from lzy_b200.utils import op_scope, tbpickle

tbpickle.install()


class MemSnapshot(Snapshot):
    def __init__(self, registry, prefix: str, blobs: Dict[str, bytes], store_counts: Dict[str, int]):
        self.registry, self.prefix, self.blobs, self.store_counts = registry, prefix, blobs, store_counts
        self.entries: Dict[str, SnapshotEntry] = {}

    def create_entry(self, name: str, typ: Type) -> SnapshotEntry:
        ser = self.registry.find_serializer_by_type(typ)
        if ser is None:
            raise TypeError(f"Cannot find serializer for type {typ}")
        e = SnapshotEntry(str(uuid.uuid4()), name, typ, ser.schema(typ), "mem")
        self.entries[e.id] = e
        return e

    def get(self, entry_id: str) -> SnapshotEntry:
        return self.entries[entry_id]

    async def put_data(self, entry_id: str, data: Any) -> None:
        e = self.entries[entry_id]
        raw = cloudpickle.dumps(data)
        md5 = hashlib.md5(raw).hexdigest()
        e.data_hash, e.storage_uri = md5, self.prefix + md5
        if e.storage_uri not in self.blobs:
            self.blobs[e.storage_uri] = raw
            self.store_counts[e.storage_uri] = self.store_counts.get(e.storage_uri, 0) + 1

    async def get_data(self, entry_id: str):
        e = self.entries[entry_id]
        try:
            uri = e.storage_uri
        except ValueError as err:
            return Absence(err)
        if uri not in self.blobs:
            return Absence(ValueError(f"Cannot find blob in storage by uri={uri}"))
        return Result(cloudpickle.loads(self.blobs[uri]))

    async def copy_data(self, from_entry_id: str, to_uri: str) -> None:
        self.blobs[to_uri] = self.blobs[self.entries[from_entry_id].storage_uri]


class RuntimeMock(Runtime):
    def __init__(self, execute: bool = True):
        self.calls: List[Any] = []
        self.execute = execute
        self.blobs: Dict[str, bytes] = {}
        self.store_counts: Dict[str, int] = {}
        self.workflow = None
        self.aborted = self.finished = 0

    async def storage(self):
        return None

    def snapshot(self, workflow, storage_uri_prefix: str):
        return MemSnapshot(workflow.owner.serializer_registry, storage_uri_prefix, self.blobs, self.store_counts)

    async def start(self, workflow) -> str:
        self.workflow = workflow
        return str(uuid.uuid4())

    async def abort(self) -> None:
        self.aborted += 1

    async def finish(self) -> None:
        self.finished += 1

    async def exec(self, calls, progress: Callable[[ProgressStep], None]) -> None:
        self.calls.extend(calls)
        if not self.execute:
            return
        snap = self.workflow.snapshot
        for c in calls:  # queue order is already a valid topological order
            async def val(eid):
                r = await snap.get_data(eid)
                assert isinstance(r, Result), f"missing input {eid}"
                return r.value
            args = [await val(e) for e in c.arg_entry_ids]
            kwargs = {k: await val(e) for k, e in c.kwarg_entry_ids.items()}
            op_scope.enter()
            try:
                res = c.signature.func.callable(*args, **kwargs)
            finally:
                op_scope.leave()
            outs = [res] if len(c.entry_ids) == 1 else list(res)
            for eid, data in zip(c.entry_ids, outs):
                self.blobs[snap.get(eid).storage_uri] = cloudpickle.dumps(data)
