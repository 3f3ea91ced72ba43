"""Snapshot: the id → data table of one workflow (pylzy/lzy/api/v1/snapshot.py:19-191).

``Snapshot`` and ``SnapshotEntry`` keep the reference's shape.  ``FrameSnapshot`` is the B200
implementation of ``DefaultSnapshot``'s contract (put_data :141-172, get_data :108-139, copy_data
:174-188) over a ``FrameStore`` that keeps every payload in an HBM slot frame:

* ``put_data`` serializes *into a frame* (tensor/ndarray: the CUDA pack kernel; anything else: cloudpickle
  bytes copied H2D), takes the frame's digest as ``data_hash`` and ``storage_uri = prefix + digest``
  — the same content-addressing rule as ``prefix + md5`` (:154-155), so equal data is stored once;
* ``get_data`` returns ``Result(value)`` or ``Absence`` exactly like the reference.
"""
from __future__ import annotations

import threading
import uuid
from abc import ABC, abstractmethod
from typing import TYPE_CHECKING, Any, Dict, Optional, Type

from ...proxy_result import Absence, Either, Result
from ...serialization.api import Schema, SerializerRegistry

if TYPE_CHECKING:
    from .b200.store import FrameStore


class _AssignedLater:
    """an attribute that exists from the start but has a value only once somebody decided it: reading it earlier
    is an error, not ``None`` (pylzy/lzy/api/v1/snapshot.py:33-53 does the same for both fields below)"""

    def __init__(self, complaint: str):
        self._complaint = complaint

    def __set_name__(self, owner, name: str) -> None:
        self._slot = "_" + name

    def __get__(self, entry, owner=None):
        if entry is None:
            return self
        value = getattr(entry, self._slot)
        if value is None:
            raise ValueError(self._complaint.format(entry=entry))
        return value

    def __set__(self, entry, value) -> None:
        setattr(entry, self._slot, value)


class SnapshotEntry:
    """one row of a workflow's data table: an argument, a result or an exception.  `storage_uri` and `data_hash`
    are decided later — for local data when it is stored (content address), for results when the call queue is
    submitted (core/workflow.py)."""
    storage_uri = _AssignedLater("Storage uri for snapshot entry with id={entry.id} is not set")
    data_hash = _AssignedLater("Data hash for snapshot entry {entry.id} is not set")

    def __init__(self, id: str, name: str, typ: Type, data_scheme: Schema, storage_name: str,
                 storage_uri: Optional[str] = None, data_hash: Optional[str] = None):
        self.id, self.name, self.typ, self.data_scheme, self.storage_name = id, name, typ, data_scheme, storage_name
        self._storage_uri, self._data_hash = storage_uri, data_hash

    def __repr__(self) -> str:
        return f"SnapshotEntry(id={self.id!r}, name={self.name!r}, typ={self.typ!r}, uri={self._storage_uri!r})"


class Snapshot(ABC):
    @abstractmethod
    def create_entry(self, name: str, typ: Type) -> SnapshotEntry:
        ...

    @abstractmethod
    async def get_data(self, entry_id: str) -> Either:
        ...

    @abstractmethod
    async def put_data(self, entry_id: str, data: Any) -> None:
        ...

    @abstractmethod
    async def copy_data(self, from_entry_id: str, to_uri: str) -> None:
        ...

    @abstractmethod
    def get(self, entry_id: str) -> SnapshotEntry:
        ...


class FrameSnapshot(Snapshot):
    # every coroutine below is a synchronous call into the frame store: LzyEventLoop may drive them on the caller's thread
    inline_coroutines = True

    def __init__(self, serializer_registry: SerializerRegistry, storage_uri: str, store: "FrameStore",
                 storage_name: str = "hbm"):
        self._registry = serializer_registry
        self._prefix = storage_uri
        self._store = store
        self._storage_name = storage_name
        self._entries: Dict[str, SnapshotEntry] = {}
        # HBM is not a disk: a payload is kept only while something can still ask for it.  An entry is
        # pinned by every queued call that consumes it and by its lazy proxy while that is alive; when
        # the last pin goes (and the data exists) the entry gives up its claim on the frame.
        self._lock = threading.RLock()
        self._pins: Dict[str, int] = {}
        self._holds: Dict[str, str] = {}      # entry id -> uri it holds a store reference on
        self._persistent: set = set()         # cache=True results stay until runtime.clear()
        self.auto_release = True

    def pin(self, entry_id: str) -> None:
        with self._lock:
            self._pins[entry_id] = self._pins.get(entry_id, 0) + 1

    def unpin(self, entry_id: str) -> None:
        with self._lock:
            n = self._pins.get(entry_id, 0) - 1
            if n > 0:
                self._pins[entry_id] = n
                return
            self._pins.pop(entry_id, None)
        self._maybe_release(entry_id)

    def persist(self, entry_id: str) -> None:
        with self._lock:
            self._persistent.add(entry_id)

    def claims(self, entry_id: str) -> int:
        """how many parties can still ask for this entry: queued consumer calls + its lazy proxy while that is alive"""
        with self._lock:
            return self._pins.get(entry_id, 0)

    def is_persistent(self, entry_id: str) -> bool:
        with self._lock:
            return entry_id in self._persistent

    def produced(self, entry_id: str) -> None:
        """an op wrote this entry: take a reference on its frame (or let it go if nobody waits for it)"""
        entry = self._entries.get(entry_id)
        if entry is None or entry._storage_uri is None:
            return
        with self._lock:
            if entry_id not in self._holds:
                self._holds[entry_id] = entry._storage_uri
                self._store.retain(entry._storage_uri)
            persistent = entry_id in self._persistent
        if persistent:
            self._store.persist(entry._storage_uri)  # cache=True: written to the spill tier (if there is one) now
        self._maybe_release(entry_id)
        self._store.enforce_budget()

    def _maybe_release(self, entry_id: str) -> None:
        if not self.auto_release:
            return
        with self._lock:
            if self._pins.get(entry_id, 0) > 0 or entry_id in self._persistent:
                return
            uri = self._holds.pop(entry_id, None)
        if uri is not None:
            self._store.release(uri)

    def create_entry(self, name: str, typ: Type) -> SnapshotEntry:
        eid = str(uuid.uuid4())
        serializer = self._registry.find_serializer_by_type(typ)
        if serializer is None:
            raise TypeError(f"Cannot find serializer for type {typ}")
        if not serializer.available():
            raise TypeError(f"Serializer for type {typ} is not available, please install {serializer.requirements()}")
        e = SnapshotEntry(eid, name, typ, serializer.schema(typ), self._storage_name)
        self._entries[eid] = e
        return e

    def get(self, entry_id: str) -> SnapshotEntry:
        return self._entries[entry_id]

    async def put_data(self, entry_id: str, data: Any) -> None:
        entry = self._entries.get(entry_id)
        if entry is None:
            raise ValueError(f"Entry with id={entry_id} does not exist")
        digest_hex = self._store.put(self._prefix, data, entry.typ, self._registry, name=entry.name)
        entry.data_hash = digest_hex
        entry.storage_uri = self._prefix + digest_hex
        with self._lock:
            if entry_id not in self._holds:
                self._holds[entry_id] = entry.storage_uri
                self._store.retain(entry.storage_uri)

    async def get_data(self, entry_id: str) -> Either:
        entry = self._entries.get(entry_id)
        if entry is None:
            raise ValueError(f"Entry with id={entry_id} does not exist")
        try:
            uri = entry.storage_uri
        except ValueError as e:
            return Absence(e)
        if not self._store.exists(uri):
            return Absence(ValueError(f"Cannot find blob in storage by uri={uri}"))
        return Result(self._store.get(uri, entry.typ, self._registry))

    async def copy_data(self, from_entry_id: str, to_uri: str) -> None:
        entry = self._entries.get(from_entry_id)
        if entry is None:
            raise ValueError(f"Entry with id={from_entry_id} does not exist")
        if not self._store.exists(entry.storage_uri):
            raise ValueError(f"Entry with id={from_entry_id} is not loaded to storage")
        self._store.alias(entry.storage_uri, to_uri)
