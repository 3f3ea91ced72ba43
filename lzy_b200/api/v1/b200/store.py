"""FrameStore — the box-local "storage" of the B200 runtime: every snapshot payload is an HBM slot
frame addressed by its storage URI.

It plays the part S3 / the file storage plays for the reference (``AsyncStorageClient``,
pylzy/lzy/storage/api.py:59-96, and the storage peers of the channel manager,
lzy/lzy-service/.../operations/graph/CreateChannels.java:75-104): data that exists before a graph
runs is a *storage producer* of its channel, and every op output "lands in storage" — except that
landing means the C++ frame store takes a reference to the producer's frame (``hbm://`` URIs,
include/lzy_b200_host.h: lzyb_frame_store_*); nothing is copied and nothing leaves HBM.

Below HBM sits an optional *spill tier* (``configure_tier``; lzyb_frame_store_set_spill / spill / load): with an
HBM budget, the least recently used frames are written to files and dropped from HBM — they stay fetchable under
the same URI, by clients and by consumers of their channels, exactly like the reference's BACKUP storage
producers (channel-manager PeerDao.java:44-53) — and ``cache=True`` results are written there as soon as they
exist, so a later process finds them under their stable URIs (pylzy/lzy/core/workflow.py:247-281) and skips the op.
"""
from __future__ import annotations

import ctypes as C
import struct
import threading
import uuid
from io import BytesIO
from typing import Any, Dict, Optional, Type

from .... import _host as H
from .... import _native as N
from ....devmem import caller_stream_scope, current_stream_ptr
from ....serialization.api import DeviceFrameSerializer, SerializerRegistry
from ....serialization.builtin import FLAG_HOST_ORIGIN, FLAG_NDARRAY, FLAG_OOB, CloudpickleSerializer

OOB_MIN_BYTES = 64 << 10  # below this an object is one plain pickle
_RING = object()  # marker: the value lives in the pinned host ring
_FLAG_SHIFT = 16  # header.dtype: low 16 bits = enum lzyb_dtype, high 16 bits = serializer flags


class _InfoView:
    """slot info with the dtype code and the serializer flags separated"""

    def __init__(self, info):
        self.dtype = int(info.dtype) & 0xFFFF
        self.flags = int(info.dtype) >> _FLAG_SHIFT
        self.ndim = int(info.ndim)
        self.shape = [int(info.shape[i]) for i in range(self.ndim)]
        self.nbytes = int(info.nbytes)
        self.gpu = int(info.gpu)


class FrameStore:
    def __init__(self, default_gpu: int = 0):
        self.default_gpu = default_gpu
        self._frames: Dict[str, N.SlotHandle] = {}
        self._inline: Dict[str, Any] = {}  # tiny control values (exception tuples) that never need a frame
        self._ring: Dict[str, int] = {}    # uri -> offset of its frame in the pinned host ring
        self._lock = threading.RLock()
        self._refs: Dict[str, int] = {}  # how many snapshot entries still need the frame at this URI
        self.bytes_put = 0
        # spill tier
        self.spill_dir: Optional[str] = None
        self.hbm_budget_bytes: int = 0
        self._durable: set = set()       # URIs whose spill file outlives this process (cache=True results)
        self._spilled_here: set = set()  # URIs this process spilled to make room: their files go when the data goes
        self._touch: Dict[str, int] = {}  # uri -> logical time of the last use (LRU order for eviction)
        self._clock = 0
        self.evictions = 0
        self.reloads = 0

    # ---- spill tier ---------------------------------------------------------------------------
    def configure_tier(self, spill_dir: Optional[str], hbm_budget_bytes: int = 0) -> None:
        """`spill_dir`: where frames go when they leave HBM (None = no tier: HBM or nothing).
        `hbm_budget_bytes`: resident bytes above which least-recently-used frames are evicted (0 = never)."""
        if hbm_budget_bytes and not spill_dir:
            raise ValueError("an HBM budget needs a spill directory: evicted frames must stay fetchable")
        H.frame_store_set_spill(spill_dir, self.default_gpu)
        self.spill_dir, self.hbm_budget_bytes = spill_dir, int(hbm_budget_bytes)

    def _touched(self, uri: str) -> None:
        self._clock += 1
        self._touch[uri] = self._clock

    def persist(self, uri: str) -> bool:
        """write the frame at `uri` to the spill tier and keep the file beyond this process"""
        if not self.spill_dir:
            return False
        with self._lock:
            self._durable.add(uri)
        if not H.frame_store_spilled(uri) and H.frame_store_resident(uri):
            H.frame_store_spill(uri)
        return H.frame_store_spilled(uri)

    def evict(self, uri: str) -> bool:
        """spill (if needed) and drop the HBM copy; the URI stays readable"""
        if not self.spill_dir or not H.frame_store_resident(uri):
            return False
        if not H.frame_store_spilled(uri):
            H.frame_store_spill(uri)
            with self._lock:
                self._spilled_here.add(uri)
        with self._lock:
            s = self._frames.pop(uri, None)
        H.frame_store_drop(uri)
        if s is not None:
            s.close()
        self.evictions += 1
        return True

    def enforce_budget(self, protect: Optional[str] = None) -> int:
        """evict least-recently-used frames (never `protect`) until the resident bytes fit the budget; returns
        how many went"""
        if not self.hbm_budget_bytes or not self.spill_dir:
            return 0
        n = 0
        while H.frame_store_bytes() > self.hbm_budget_bytes:
            with self._lock:
                known = (set(self._refs) | set(self._frames)) - {protect}
                victims = sorted((u for u in known if u not in self._inline), key=lambda u: self._touch.get(u, 0))
            victim = next((u for u in victims if H.frame_store_resident(u)), None)
            if victim is None or not self.evict(victim):
                break
            n += 1
        return n

    # ---- writing ------------------------------------------------------------------------------
    def pack_object(self, data: Any, typ: Type, registry: SerializerRegistry, slot_factory, stream: int = 0):
        """Serialize `data` into a fresh frame obtained from `slot_factory(nbytes) -> SlotHandle`.
        Returns the slot (not yet committed)."""
        ser = registry.find_serializer_by_instance(data)
        if isinstance(ser, DeviceFrameSerializer) and ser.can_frame(data):
            code, shape, nbytes, flags = ser.frame_plan(data)
            slot = slot_factory(nbytes, getattr(data, "device", None))
            slot.set_meta(code | (flags << _FLAG_SHIFT), shape)
            ser.pack(data, slot, stream)
            return slot
        ser = registry.find_serializer_by_type(typ) or ser
        if ser is None or isinstance(ser, DeviceFrameSerializer):
            ser = registry.find_serializer_by_data_format("pickle")
        if isinstance(ser, CloudpickleSerializer):
            skeleton, bufs = ser.dumps_oob(data)
            if bufs and sum(b.nbytes for b in bufs) >= OOB_MIN_BYTES:
                return self._pack_composite(skeleton, bufs, slot_factory, stream)
            raw = skeleton if not bufs else None
        else:
            raw = None
        if raw is None:
            buf = BytesIO()
            ser.serialize(data, buf)
            raw = buf.getvalue()
        slot = slot_factory(len(raw), None)
        slot.set_meta(N.DT_PICKLE | (FLAG_HOST_ORIGIN << _FLAG_SHIFT), (len(raw),))
        slot.pack_host(C.cast(C.c_char_p(raw), C.c_void_p).value or 0, len(raw), stream)
        slot._keepalive = raw
        return slot

    @staticmethod
    def _pack_composite(skeleton: bytes, bufs, slot_factory, stream: int):
        """one frame = [table][skeleton][buffer 0][buffer 1]…, every part 16-byte aligned; the table is
        (n, skeleton_len, (offset, length) × n) as little-endian u64"""
        pad = lambda n: (n + 15) & ~15
        table_len = pad(8 * (2 + 2 * len(bufs)))
        off = table_len + pad(len(skeleton))
        entries = []
        for b in bufs:
            entries.append((off, b.nbytes))
            off += pad(b.nbytes)
        total = off
        head = struct.pack(f"<{2 + 2 * len(bufs)}Q", len(bufs), len(skeleton), *[x for e in entries for x in e])
        head = head.ljust(table_len, b"\0") + skeleton
        slot = slot_factory(total, None)
        slot.set_meta(N.DT_PICKLE | ((FLAG_HOST_ORIGIN | FLAG_OOB) << _FLAG_SHIFT), (total,))
        keep = [head]
        slot.pack_host_part(0, C.cast(C.c_char_p(head), C.c_void_p).value or 0, len(head), False, stream)
        for i, (b, (o, n)) in enumerate(zip(bufs, entries)):
            arr = (C.c_char * n).from_buffer(b) if not b.readonly else (C.c_char * n).from_buffer_copy(b)
            keep.append(arr)
            slot.pack_host_part(o, C.addressof(arr), n, i == len(bufs) - 1, stream)
        slot._keepalive = (keep, bufs)
        return slot

    @staticmethod
    def _unpack_composite(slot: N.SlotHandle, nbytes: int):
        host = bytearray(nbytes)
        view = (C.c_char * nbytes).from_buffer(host)
        slot.copy_to_host(C.addressof(view), nbytes)
        del view
        n, skel_len = struct.unpack_from("<2Q", host, 0)
        table = struct.unpack_from(f"<{2 * n}Q", host, 16)
        table_len = (8 * (2 + 2 * n) + 15) & ~15
        mv = memoryview(host)
        buffers = [mv[table[2 * i]: table[2 * i] + table[2 * i + 1]] for i in range(n)]
        return CloudpickleSerializer.loads_oob(mv[table_len: table_len + skel_len], buffers)

    def put(self, prefix: str, data: Any, typ: Type, registry: SerializerRegistry, name: str = "") -> str:
        def factory(nbytes, device):
            gpu = device.index if device is not None and getattr(device, "type", "") == "cuda" else self.default_gpu
            return N.SlotHandle.create_output(f"put/{name}/{uuid.uuid4()}", gpu, nbytes)

        probe_gpu = getattr(getattr(data, "device", None), "index", None)
        gpu = probe_gpu if probe_gpu is not None else self.default_gpu
        with caller_stream_scope(gpu):  # the caller filled `data` on its own stream: pack in that stream's order
            stream = current_stream_ptr(gpu)
            slot = self.pack_object(data, typ, registry, factory, stream)
            digest = slot.commit(stream)  # waits for the pack: `data` may be reused as soon as put returns
        slot._keepalive = None
        # the address covers how the bytes are to be read, not only the bytes: equal payloads of another dtype, shape
        # or serializer are different values (the reference's md5 runs over the serialized stream, snapshot.py:147-155)
        info = slot.info()
        hexd = N.digest_hex(N.content_key(digest, int(info.dtype), [int(info.shape[i]) for i in range(int(info.ndim))]))
        uri = prefix + hexd
        with self._lock:
            if uri in self._frames:  # equal data is stored once (snapshot.py:157-170)
                slot.close()
            else:
                self._frames[uri] = slot
                H.frame_store_put(uri, slot.h)
                self.bytes_put += slot.nbytes
            self._touched(uri)
        self.enforce_budget()
        return hexd

    def put_inline(self, uri: str, value: Any) -> None:
        """Small host-side control objects (an op's exception tuple) never need HBM: they are pickled,
        framed into the pinned host ring by the CUDA frame kernel (header + digest written into the
        ring, `lzyb_frame_host_object`) and verified by a scan when read back."""
        if value is None:
            with self._lock:
                self._inline[uri] = None
            return
        import cloudpickle
        off = N.frame_host_object(cloudpickle.dumps(value))
        with self._lock:
            old = self._ring.pop(uri, None)
            self._ring[uri] = off
            self._inline[uri] = _RING
        if old is not None:
            N.ring_release(old)

    # ---- reading ------------------------------------------------------------------------------
    def _slot(self, uri: str) -> Optional[N.SlotHandle]:
        with self._lock:
            self._touched(uri)
            s = self._frames.get(uri)
            if s is not None:
                return s
            was_resident = H.frame_store_resident(uri)
            h = H.frame_store_get(uri)  # produced by an op (the slots runtime registered it), or back from the spill tier
            if h is None:
                return None
            if not was_resident:
                self.reloads += 1
            s = N.SlotHandle.adopt(h, owned=True)
            self._frames[uri] = s
            return s

    def exists(self, uri: str) -> bool:
        with self._lock:
            if uri in self._inline or uri in self._frames:
                return True
        # resident in the C++ store (an op produced it) or waiting in the spill tier; asking does not load it
        return H.frame_store_resident(uri) or (self.spill_dir is not None and H.frame_store_spilled(uri))

    def unpack_slot(self, slot: N.SlotHandle, typ: Type, registry: SerializerRegistry, private: bool, stream: int = 0) -> Any:
        """`stream`: the CUDA stream the value will be used on (0 = the calling thread's current stream on the slot's
        GPU).  A zero-copy view is only as good as its ordering: that stream is made to wait for the frame's last writer
        (a pack that may still be running when the slot was committed in stream order, lzyb_set_lazy_commit)."""
        info = _InfoView(slot.meta())
        if info.dtype == N.DT_PICKLE and info.flags & FLAG_OOB:
            return self._unpack_composite(slot, info.nbytes)
        if info.dtype == N.DT_PICKLE:
            raw = C.create_string_buffer(max(1, info.nbytes))
            if info.nbytes:
                slot.copy_to_host(C.addressof(raw), info.nbytes)
            ser = registry.find_serializer_by_type(typ)
            if ser is None or isinstance(ser, DeviceFrameSerializer):
                ser = registry.find_serializer_by_data_format("pickle")
            return ser.deserialize(BytesIO(raw.raw[:info.nbytes]), typ if isinstance(typ, type) else None)
        fmt = "lzyb_ndarray_frame" if info.flags & FLAG_NDARRAY else "lzyb_tensor_frame"
        ser = registry.find_serializer_by_data_format(fmt)
        if not info.flags & (FLAG_NDARRAY | FLAG_HOST_ORIGIN):  # a device view: the host copies order themselves inside the C ABI
            slot.wait(stream or current_stream_ptr(slot.gpu))
        value = ser.unpack(slot, info, private)
        if not private and hasattr(value, "is_cuda") and value.is_cuda:
            value = value.clone()  # a client read gets its own copy, like a fresh deserialization
        return value

    def get(self, uri: str, typ: Type, registry: SerializerRegistry) -> Any:
        v = None
        with self._lock:
            if uri in self._inline:
                v = self._inline[uri]
                if v is not _RING:
                    return v
                off = self._ring[uri]
        if v is _RING:
            import cloudpickle
            return cloudpickle.loads(N.ring_read(off))
        reloads = self.reloads
        slot = self._slot(uri)
        if slot is None:
            raise ValueError(f"Cannot find blob in storage by uri={uri}")
        if self.reloads != reloads:
            self.enforce_budget(protect=uri)  # what came back from the tier pushes something colder out
        with caller_stream_scope(slot.gpu):  # the caller will use the value on its own stream
            value = self.unpack_slot(slot, typ, registry, private=False)
            if getattr(value, "is_cuda", False):
                # the copy must have left the frame before the frame can be released (clear / eviction)
                import torch
                torch.cuda.current_stream(slot.gpu).synchronize()
        return value

    def alias(self, from_uri: str, to_uri: str) -> None:
        slot = self._slot(from_uri)
        if slot is None:
            raise ValueError(f"no data at {from_uri}")
        H.frame_store_put(to_uri, slot.h)

    def gpu_of(self, uri: str) -> Optional[int]:
        s = self._slot(uri)
        return None if s is None else s.gpu

    def nbytes_of(self, uri: str) -> int:
        s = self._slot(uri)
        return 0 if s is None else s.nbytes

    # ---- lifetime -----------------------------------------------------------------------------
    def retain(self, uri: str) -> None:
        with self._lock:
            self._refs[uri] = self._refs.get(uri, 0) + 1

    def release(self, uri: str) -> None:
        """drop one entry's claim on `uri`; the frame goes back to the HBM pool with the last one"""
        with self._lock:
            n = self._refs.get(uri, 0) - 1
            if n > 0:
                self._refs[uri] = n
                return
            self._refs.pop(uri, None)
        self.drop(uri)

    def drop(self, uri: str) -> None:
        with self._lock:
            self._inline.pop(uri, None)
            off = self._ring.pop(uri, None)
            s = self._frames.pop(uri, None)
            self._touch.pop(uri, None)
            scratch = uri in self._spilled_here and uri not in self._durable
            self._spilled_here.discard(uri)
        if off is not None:
            N.ring_release(off)
        H.frame_store_drop(uri)
        if s is not None:
            s.close()
        if scratch:
            H.frame_store_unspill(uri)  # an evicted frame nobody can ask for any more; durable files stay

    def clear(self, prefix: str = "hbm://") -> None:
        """drop every frame under `prefix`, including op outputs nobody has read yet"""
        with self._lock:
            uris = [u for u in set(self._frames) | set(self._inline) | set(self._spilled_here) if u.startswith(prefix)]
            for u in [u for u in self._refs if u.startswith(prefix)]:
                self._refs.pop(u, None)
        for u in uris:
            self.drop(u)
        H.frame_store_clear(prefix)

    @staticmethod
    def resident_bytes() -> int:
        return H.frame_store_bytes()

    def __len__(self):
        return len(self._frames)
