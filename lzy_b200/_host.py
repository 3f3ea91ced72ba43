"""ctypes binding of the host-side control logic in liblzy_b200.so (include/lzy_b200_host.h):
the channel manager's matcher and the graph dispatcher.  No CUDA involved."""
from __future__ import annotations

import ctypes as C
import json
from typing import List, Optional, Sequence, Tuple

from . import _native as N

lib = N.lib
_H = C.c_uint64

PEER_SLOT, PEER_STORAGE = 1, 2
ROLE_PRODUCER, ROLE_CONSUMER = 1, 2
TR_PENDING, TR_ACTIVE, TR_COMPLETED, TR_FAILED = 0, 1, 2, 3
PRIO_PRIMARY, PRIO_BACKUP = 7, 5
ACT_START_TRANSFER, ACT_ABORT_WORKFLOW = 1, 2
G_WAITING, G_EXECUTING, G_COMPLETED, G_FAILED = 0, 1, 2, 3
T_WAITING, T_EXECUTING, T_SUCCESS, T_ERROR = 0, 1, 2, 3


class PeerDesc(C.Structure):
    _fields_ = [("peer_id", C.c_char * 128), ("kind", C.c_int32), ("reserved", C.c_int32), ("url", C.c_char * 512)]

    def as_tuple(self):
        return self.peer_id.decode(), int(self.kind), self.url.decode()


class BindResult(C.Structure):
    _fields_ = [("has_peer", C.c_int32), ("reserved", C.c_int32), ("peer", PeerDesc), ("transfer_id", C.c_char * 64)]


class CmAction(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("channel_id", C.c_char * 64),
                ("transfer_id", C.c_char * 64), ("consumer_id", C.c_char * 128), ("consumer_url", C.c_char * 512),
                ("producer", PeerDesc), ("dial_url", C.c_char * 512), ("execution_id", C.c_char * 128),
                ("workflow_name", C.c_char * 128), ("reason", C.c_char * 256), ("idempotency_key", C.c_char * 128)]


_s, _i = C.c_char_p, C.c_int
_SIGS = {
    "lzyb_cm_create": (_i, [C.POINTER(_H)]),
    "lzyb_cm_free": (_i, [_H]),
    "lzyb_cm_seed": (_i, [_H, C.c_uint64]),
    "lzyb_cm_get_or_create": (_i, [_H, _s, _s, _s, _s, _s, _s, _i, _s, C.c_char * 64]),
    "lzyb_cm_destroy": (_i, [_H, _s, _s]),
    "lzyb_cm_destroy_all": (_i, [_H, _s, _s]),
    "lzyb_cm_bind": (_i, [_H, _s, _s, _s, _s, _s, _i, C.POINTER(BindResult)]),
    "lzyb_cm_unbind": (_i, [_H, _s, _s]),
    "lzyb_cm_transfer_completed": (_i, [_H, _s, _s, _s]),
    "lzyb_cm_transfer_failed": (_i, [_H, _s, _s, _s, _s, C.POINTER(BindResult)]),
    "lzyb_cm_channels_status": (C.c_int64, [_H, _s, C.POINTER(_s), _i, C.c_char_p, C.c_uint64]),
    "lzyb_cm_poll_action": (_i, [_H, C.POINTER(CmAction), _i]),
    "lzyb_cm_action_result": (_i, [_H, _s, _s, _i, _s]),
    "lzyb_cm_restore_actions": (_i, [_H]),
    "lzyb_cm_transfer_state": (_i, [_H, _s, _s, C.POINTER(_i)]),
    "lzyb_cm_peer_priority": (_i, [_H, _s, _s, C.POINTER(_i)]),
    "lzyb_cm_is_aborted": (_i, [_H, _s, C.c_char_p, C.c_uint64]),
    "lzyb_slots_set_backend": (_i, [C.c_void_p]),
    "lzyb_slots_create": (_i, [_H, _i, _s, C.POINTER(_H)]),
    "lzyb_slots_free": (_i, [_H]),
    "lzyb_slots_set_chunk_bytes": (_i, [_H, C.c_uint64]),
    "lzyb_slots_set_peer_chunk_bytes": (_i, [_H, C.c_uint64]),
    "lzyb_ctx_create": (_i, [_H, _s, _s, C.POINTER(_H)]),
    "lzyb_ctx_add_slot": (_i, [_H, _s, _s, _i]),
    "lzyb_ctx_add_slot_flags": (_i, [_H, _s, _s, _i, C.c_uint32]),
    "lzyb_ctx_before_execution": (_i, [_H, _i]),
    "lzyb_ctx_prefetch": (_i, [_H]),
    "lzyb_ctx_input_handle": (_i, [_H, _s, C.POINTER(_H)]),
    "lzyb_ctx_output_open": (_i, [_H, _s, C.c_uint64, C.POINTER(_H)]),
    "lzyb_ctx_output_done": (_i, [_H, _s]),
    "lzyb_ctx_after_execution": (_i, [_H, _i]),
    "lzyb_ctx_close": (_i, [_H]),
    "lzyb_ctx_input_state": (_i, [_H, _s]),
    "lzyb_slots_start_transfer": (_i, [_H, _s, C.POINTER(PeerDesc), _s]),
    "lzyb_slots_read": (C.c_int64, [_H, _s, C.c_uint64, C.c_void_p, C.c_uint64]),
    "lzyb_cm_pump_start": (_i, [_H]),
    "lzyb_cm_pump_stop": (_i, [_H]),
    "lzyb_blob_put": (_i, [_s, C.c_void_p, C.c_uint64]),
    "lzyb_blob_size": (C.c_int64, [_s]),
    "lzyb_blob_get": (C.c_int64, [_s, C.c_void_p, C.c_uint64]),
    "lzyb_blob_delete": (_i, [_s]),
    "lzyb_slots_threads_spawned": (C.c_uint64, []),
    "lzyb_slots_set_cm_client": (_i, [C.c_void_p]),
    "lzyb_slots_set_peer_resolver": (_i, [C.c_void_p]),
    "lzyb_slots_describe": (_i, [_H, _s, C.c_void_p]),
    "lzyb_cm_serve_local": (_i, [_H, _s]),
    "lzyb_cm_stop_local": (_i, [_H]),
    "lzyb_slots_connect_cm_local": (_i, [_s, C.POINTER(_H)]),
    "lzyb_slots_disconnect_cm_local": (_i, []),
    "lzyb_slots_serve_local": (_i, [_H, _s]),
    "lzyb_slots_stop_local": (_i, [_H]),
    "lzyb_slots_transfer_log": (C.c_int64, [_H, C.c_char_p, C.c_uint64, C.c_int]),
    "lzyb_frame_store_put": (_i, [_s, _H]),
    "lzyb_frame_store_get": (_i, [_s, C.POINTER(_H)]),
    "lzyb_frame_store_drop": (_i, [_s]),
    "lzyb_frame_store_clear": (_i, [_s]),
    "lzyb_frame_store_bytes": (C.c_int64, []),
    "lzyb_frame_store_set_spill": (_i, [_s, C.c_int]),
    "lzyb_frame_store_spill": (_i, [_s]),
    "lzyb_frame_store_load": (_i, [_s, C.c_int, C.POINTER(_H)]),
    "lzyb_frame_store_spilled": (_i, [_s]),
    "lzyb_frame_store_resident": (_i, [_s]),
    "lzyb_frame_store_unspill": (_i, [_s]),
    "lzyb_slots_set_republish": (_i, [_H, _i]),
    "lzyb_graph_create": (_i, [_i, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i, _i, C.POINTER(_H)]),
    "lzyb_graph_free": (_i, [_H]),
    "lzyb_graph_next": (_i, [_H, C.POINTER(C.c_int32), _i]),
    "lzyb_graph_task_done": (_i, [_H, C.c_int32, _i, _s]),
    "lzyb_graph_executing": (_i, [_H, C.POINTER(C.c_int32), _i]),
    "lzyb_graph_status": (_i, [_H]),
    "lzyb_graph_task_status": (_i, [_H, C.c_int32]),
    "lzyb_graph_stop": (_i, [_H, _s, C.POINTER(C.c_int32), _i]),
    "lzyb_graph_error": (_i, [_H, C.c_char_p, C.c_uint64, C.POINTER(C.c_int32)]),
    "lzyb_graph_topo_order": (_i, [_H, C.POINTER(C.c_int32), _i]),
    "lzyb_graph_has_cycle": (_i, [_H]),
}
for _n, (_r, _a) in _SIGS.items():
    _f = getattr(lib, _n)
    _f.restype, _f.argtypes = _r, _a
EXPORTED_SYMBOLS = tuple(_SIGS)


def _b(s: Optional[str]) -> Optional[bytes]:
    return None if s is None else s.encode()


class GrpcStatusError(N.LzybError):
    """Control-plane failure; ``status`` is the gRPC status name the reference would return."""
    NAMES = {N.E_NOT_FOUND: "NOT_FOUND", N.E_PRECONDITION: "FAILED_PRECONDITION", N.E_UNAVAILABLE: "UNAVAILABLE",
             N.E_INTERNAL: "INTERNAL", N.E_UNIMPLEMENTED: "UNIMPLEMENTED", N.E_INVALID: "INVALID_ARGUMENT",
             N.E_EXISTS: "ALREADY_EXISTS"}

    def __init__(self, code, message):
        super().__init__(code, message)
        self.status = self.NAMES.get(code, "UNKNOWN")


def _check(rc: int) -> int:
    if rc >= 0:
        return rc
    msg = N.last_error()
    if rc in GrpcStatusError.NAMES:
        raise GrpcStatusError(rc, msg)
    raise N.LzybError(rc, msg)


class Peer:
    __slots__ = ("peer_id", "kind", "url")

    def __init__(self, peer_id: str, kind: int, url: str):
        self.peer_id, self.kind, self.url = peer_id, kind, url

    @property
    def is_storage(self) -> bool:
        return self.kind == PEER_STORAGE

    def __repr__(self):
        return f"Peer({self.peer_id!r}, {'storage' if self.is_storage else 'slot'}, {self.url!r})"


class ChannelManager:
    """In-process LzyChannelManager + LzyChannelManagerPrivate
    (lzy/channel-manager-api/src/main/proto/ai/lzy/v1/channel/channel-manager*.proto)."""

    def __init__(self, seed: Optional[int] = None):
        h = _H()
        _check(lib.lzyb_cm_create(C.byref(h)))
        self.h = h.value
        if seed is not None:
            _check(lib.lzyb_cm_seed(self.h, seed))

    def close(self):
        if self.h:
            lib.lzyb_cm_stop_local(self.h)
            lib.lzyb_cm_free(self.h)
            self.h = 0

    # -- box-local transport (csrc/host/local_rpc.cpp) -----------------------------------------------
    def serve_local(self, socket_path: str) -> None:
        """Bind / Unbind / TransferCompleted / TransferFailed of this matcher on a unix socket, served by library threads:
        what the worker processes of this box dial instead of the gRPC port (`connect_cm_local`)"""
        _check(lib.lzyb_cm_serve_local(self.h, _b(socket_path)))

    def stop_local(self) -> None:
        _check(lib.lzyb_cm_stop_local(self.h))

    # -- private API ---------------------------------------------------------------------------
    def get_or_create(self, user_id: str, workflow_name: str, execution_id: str, *, producer_uri: Optional[str] = None,
                      consumer_uri: Optional[str] = None, data_format: str = "", scheme_format: str = "",
                      scheme_content: str = "") -> str:
        role = ROLE_PRODUCER if producer_uri is not None else (ROLE_CONSUMER if consumer_uri is not None else 0)
        uri = producer_uri if producer_uri is not None else consumer_uri
        out = (C.c_char * 64)()
        _check(lib.lzyb_cm_get_or_create(self.h, _b(user_id), _b(workflow_name), _b(execution_id), _b(data_format),
                                         _b(scheme_format), _b(scheme_content), role, _b(uri), out))
        return out.value.decode()

    def destroy(self, channel_id: str, reason: str = "") -> None:
        _check(lib.lzyb_cm_destroy(self.h, _b(channel_id), _b(reason)))

    def destroy_all(self, execution_id: str, reason: str = "") -> None:
        _check(lib.lzyb_cm_destroy_all(self.h, _b(execution_id), _b(reason)))

    # -- public API ----------------------------------------------------------------------------
    def bind(self, channel_id: str, role: int, peer_id: str, peer_url: str = "", execution_id: str = "",
             idempotency_key: Optional[str] = None) -> Tuple[Optional[Peer], Optional[str]]:
        r = BindResult()
        _check(lib.lzyb_cm_bind(self.h, _b(idempotency_key), _b(channel_id), _b(peer_id), _b(execution_id),
                                _b(peer_url), role, C.byref(r)))
        if not r.has_peer:
            return None, None
        return Peer(*r.peer.as_tuple()), r.transfer_id.decode()

    def unbind(self, channel_id: str, peer_id: str) -> None:
        _check(lib.lzyb_cm_unbind(self.h, _b(channel_id), _b(peer_id)))

    def transfer_completed(self, channel_id: str, transfer_id: str, idempotency_key: Optional[str] = None) -> None:
        _check(lib.lzyb_cm_transfer_completed(self.h, _b(idempotency_key), _b(channel_id), _b(transfer_id)))

    def transfer_failed(self, channel_id: str, transfer_id: str, description: str = "",
                        idempotency_key: Optional[str] = None) -> Tuple[Peer, str]:
        r = BindResult()
        _check(lib.lzyb_cm_transfer_failed(self.h, _b(idempotency_key), _b(channel_id), _b(transfer_id),
                                           _b(description), C.byref(r)))
        return Peer(*r.peer.as_tuple()), r.transfer_id.decode()

    def channels_status(self, execution_id: str = "", channel_ids: Sequence[str] = ()) -> List[dict]:
        arr = (C.c_char_p * max(1, len(channel_ids)))(*[c.encode() for c in channel_ids])
        cap = 4096
        while True:  # the document may grow between a sizing call and the real one (peers bind concurrently): ask until it fits
            buf = C.create_string_buffer(cap)
            need = int(_check(lib.lzyb_cm_channels_status(self.h, _b(execution_id), arr, len(channel_ids), buf, cap)))
            if need <= cap:
                return json.loads(buf.value.decode())["channels"]
            cap = need + 1024

    # -- deferred actions ------------------------------------------------------------------------
    def poll_action(self, timeout_ms: int = 0) -> Optional[dict]:
        a = CmAction()
        if not _check(lib.lzyb_cm_poll_action(self.h, C.byref(a), timeout_ms)):
            return None
        return {"kind": int(a.kind), "channel_id": a.channel_id.decode(), "transfer_id": a.transfer_id.decode(),
                "consumer_id": a.consumer_id.decode(), "consumer_url": a.consumer_url.decode(),
                "producer": Peer(*a.producer.as_tuple()), "dial_url": a.dial_url.decode(),
                "execution_id": a.execution_id.decode(), "workflow_name": a.workflow_name.decode(),
                "reason": a.reason.decode(), "idempotency_key": a.idempotency_key.decode()}

    def action_result(self, channel_id: str, transfer_id: str, ok: bool, reason: str = "") -> None:
        _check(lib.lzyb_cm_action_result(self.h, _b(channel_id), _b(transfer_id), int(ok), _b(reason)))

    def restore_actions(self) -> int:
        return _check(lib.lzyb_cm_restore_actions(self.h))

    # -- introspection ---------------------------------------------------------------------------
    def transfer_state(self, channel_id: str, transfer_id: str) -> int:
        st = C.c_int()
        _check(lib.lzyb_cm_transfer_state(self.h, _b(channel_id), _b(transfer_id), C.byref(st)))
        return st.value

    def peer_priority(self, channel_id: str, peer_id: str) -> int:
        p = C.c_int()
        _check(lib.lzyb_cm_peer_priority(self.h, _b(channel_id), _b(peer_id), C.byref(p)))
        return p.value

    def aborted(self, execution_id: str) -> Optional[str]:
        buf = C.create_string_buffer(256)
        if _check(lib.lzyb_cm_is_aborted(self.h, _b(execution_id), buf, 256)):
            return buf.value.decode()
        return None


class GraphDispatcher:
    """Dependency-count dispatcher with the BFS processor's observable behaviour."""

    def __init__(self, n_tasks: int, edges: Sequence[Tuple[int, int]], allow_cycles: bool = True):
        fr = (C.c_int32 * max(1, len(edges)))(*[e[0] for e in edges])
        to = (C.c_int32 * max(1, len(edges)))(*[e[1] for e in edges])
        h = _H()
        _check(lib.lzyb_graph_create(n_tasks, fr, to, len(edges), int(allow_cycles), C.byref(h)))
        self.h, self.n = h.value, n_tasks
        self._buf = (C.c_int32 * max(1, n_tasks))()

    def close(self):
        if self.h:
            lib.lzyb_graph_free(self.h)
            self.h = 0

    def next(self) -> List[int]:
        k = _check(lib.lzyb_graph_next(self.h, self._buf, self.n))
        return [int(self._buf[i]) for i in range(k)]

    def task_done(self, task: int, ok: bool = True, error: str = "") -> None:
        _check(lib.lzyb_graph_task_done(self.h, task, int(ok), _b(error)))

    def executing(self) -> List[int]:
        k = _check(lib.lzyb_graph_executing(self.h, self._buf, self.n))
        return [int(self._buf[i]) for i in range(k)]

    @property
    def status(self) -> int:
        return _check(lib.lzyb_graph_status(self.h))

    def task_status(self, task: int) -> int:
        return _check(lib.lzyb_graph_task_status(self.h, task))

    def stop(self, description: str = "") -> List[int]:
        k = _check(lib.lzyb_graph_stop(self.h, _b(description), self._buf, self.n))
        return [int(self._buf[i]) for i in range(k)]

    def error(self) -> Tuple[bool, str, int]:
        buf = C.create_string_buffer(512)
        t = C.c_int32(-1)
        failed = _check(lib.lzyb_graph_error(self.h, buf, 512, C.byref(t)))
        return bool(failed), buf.value.decode(), int(t.value)

    def topo_order(self) -> List[int]:
        k = _check(lib.lzyb_graph_topo_order(self.h, self._buf, self.n))
        return [int(self._buf[i]) for i in range(k)]

    @property
    def has_cycle(self) -> bool:
        return bool(_check(lib.lzyb_graph_has_cycle(self.h)))


DIR_INPUT, DIR_OUTPUT = 1, 2
IS_BINDING, IS_WAITING_FOR_PEER, IS_DOWNLOADING, IS_READY, IS_CLOSED = 0, 1, 2, 3, 4


def blob_put(uri: str, data: bytes) -> None:
    _check(lib.lzyb_blob_put(_b(uri), data, len(data)))


def blob_get(uri: str) -> bytes:
    n = _check(lib.lzyb_blob_size(_b(uri)))
    buf = C.create_string_buffer(max(1, n))
    _check(lib.lzyb_blob_get(_b(uri), buf, n))
    return buf.raw[:n]


def blob_exists(uri: str) -> bool:
    return lib.lzyb_blob_size(_b(uri)) >= 0


def blob_delete(uri: str) -> None:
    _check(lib.lzyb_blob_delete(_b(uri)))


def frame_store_put(uri: str, slot_handle: int) -> None:
    _check(lib.lzyb_frame_store_put(_b(uri), slot_handle))


def frame_store_get(uri: str) -> Optional[int]:
    h = _H()
    rc = lib.lzyb_frame_store_get(_b(uri), C.byref(h))
    if rc == N.E_NOT_FOUND:
        return None
    _check(rc)
    return h.value


def frame_store_drop(uri: str) -> None:
    _check(lib.lzyb_frame_store_drop(_b(uri)))


def frame_store_clear(prefix: str = "") -> int:
    return _check(lib.lzyb_frame_store_clear(_b(prefix)))


def frame_store_bytes() -> int:
    return int(lib.lzyb_frame_store_bytes())


def frame_store_set_spill(directory: Optional[str], default_gpu: int = 0) -> None:
    """the tier below HBM residency: spilled frames live as files under `directory` (None switches it off)"""
    _check(lib.lzyb_frame_store_set_spill(_b(directory or ""), default_gpu))


def frame_store_spill(uri: str) -> None:
    _check(lib.lzyb_frame_store_spill(_b(uri)))


def frame_store_load(uri: str, gpu: int) -> Optional[int]:
    h = _H()
    rc = lib.lzyb_frame_store_load(_b(uri), gpu, C.byref(h))
    if rc == N.E_NOT_FOUND:
        return None
    _check(rc)
    return h.value


def frame_store_spilled(uri: str) -> bool:
    return _check(lib.lzyb_frame_store_spilled(_b(uri))) == 1


def frame_store_resident(uri: str) -> bool:
    return _check(lib.lzyb_frame_store_resident(_b(uri))) == 1


def frame_store_unspill(uri: str) -> None:
    _check(lib.lzyb_frame_store_unspill(_b(uri)))


SLOTF_SOLE_CONSUMER = 1


class ExecutionContext:
    """SlotsExecutionContext (lzy/slots/src/main/java/ai/lzy/slots/SlotsExecutionContext.java:45-107)."""

    def __init__(self, slots: "Slots", execution_id: str, task_id: str):
        h = _H()
        _check(lib.lzyb_ctx_create(slots.h, _b(execution_id), _b(task_id), C.byref(h)))
        self.h, self.slots = h.value, slots

    def add_slot(self, name: str, channel_id: str, direction: int, sole_consumer: bool = False) -> None:
        """`sole_consumer` (outputs): exactly one consumer, on this GPU, and no other reader — it takes the frame over"""
        _check(lib.lzyb_ctx_add_slot_flags(self.h, _b(name), _b(channel_id), direction, SLOTF_SOLE_CONSUMER if sole_consumer else 0))

    def prefetch(self) -> None:
        _check(lib.lzyb_ctx_prefetch(self.h))

    def before_execution(self, timeout_ms: int = -1) -> None:
        _check(lib.lzyb_ctx_before_execution(self.h, timeout_ms))

    def input_handle(self, name: str) -> int:
        h = _H()
        _check(lib.lzyb_ctx_input_handle(self.h, _b(name), C.byref(h)))
        return h.value

    def input_state(self, name: str) -> int:
        return _check(lib.lzyb_ctx_input_state(self.h, _b(name)))

    def output_open(self, name: str, nbytes: int) -> int:
        h = _H()
        _check(lib.lzyb_ctx_output_open(self.h, _b(name), nbytes, C.byref(h)))
        return h.value

    def output_done(self, name: str) -> None:
        _check(lib.lzyb_ctx_output_done(self.h, _b(name)))

    def after_execution(self, timeout_ms: int = -1) -> None:
        _check(lib.lzyb_ctx_after_execution(self.h, timeout_ms))

    def close(self) -> None:
        if self.h:
            lib.lzyb_ctx_close(self.h)
            self.h = 0


class Slots:
    """One per worker: Slots (lzy/slots/src/main/java/ai/lzy/slots/Slots.java:44-106) with the
    LzySlotsApi entry points (StartTransfer / Read) as methods."""

    def __init__(self, cm: ChannelManager, gpu: int, api_url: str):
        h = _H()
        _check(lib.lzyb_slots_create(cm.h, gpu, _b(api_url), C.byref(h)))
        self.h, self.cm, self.gpu, self.api_url = h.value, cm, gpu, api_url

    def set_chunk_bytes(self, n: int) -> None:
        _check(lib.lzyb_slots_set_chunk_bytes(self.h, n))

    def set_peer_chunk_bytes(self, n: int) -> None:
        """slice size of transfers whose source is on another GPU (0 = one launch)"""
        _check(lib.lzyb_slots_set_peer_chunk_bytes(self.h, n))

    def transfer_log(self, clear: bool = False) -> List[dict]:
        """one record per finished input transfer: bytes, ns, final peer, failovers, outcome"""
        import json
        need = _check(lib.lzyb_slots_transfer_log(self.h, None, 0, 0))
        while True:
            buf = C.create_string_buffer(int(need) + 4096)
            got = _check(lib.lzyb_slots_transfer_log(self.h, buf, len(buf), int(clear)))
            if got <= len(buf):
                return json.loads(buf.value.decode())
            need = got

    def describe(self, peer_id: str) -> bytes:
        """the 256-byte descriptor of a producer slot of this worker, for a consumer in another process"""
        desc = N.IpcDesc()
        _check(lib.lzyb_slots_describe(self.h, _b(peer_id), C.byref(desc)))
        return bytes(desc)

    def set_republish(self, on: bool) -> None:
        _check(lib.lzyb_slots_set_republish(self.h, int(on)))

    def context(self, execution_id: str, task_id: str) -> ExecutionContext:
        return ExecutionContext(self, execution_id, task_id)

    def start_transfer(self, slot_id: str, peer: Peer, transfer_id: str) -> None:
        d = PeerDesc()
        d.peer_id, d.kind, d.url = peer.peer_id.encode(), peer.kind, peer.url.encode()
        _check(lib.lzyb_slots_start_transfer(self.h, _b(slot_id), C.byref(d), _b(transfer_id)))

    def read(self, peer_id: str, offset: int, max_len: int = 2 << 20) -> Optional[bytes]:
        """one ReadDataChunk; None = the EOS control message"""
        buf = C.create_string_buffer(max_len)
        n = lib.lzyb_slots_read(self.h, _b(peer_id), offset, buf, max_len)
        if n == N.EOF:
            return None
        _check(n)
        return buf.raw[:n]

    def serve_local(self, socket_path: str) -> None:
        """LzybPeerApi.Describe of this worker on a unix socket (library threads); announce it as "uds=<path>" in the
        query of `api_url` and consumers of this box resolve this worker's peers there"""
        _check(lib.lzyb_slots_serve_local(self.h, _b(socket_path)))

    def close(self) -> None:
        if self.h:
            lib.lzyb_slots_stop_local(self.h)
            lib.lzyb_slots_free(self.h)
            self.h = 0


class LocalChannelManager:
    """stands where a `ChannelManager` would in `Slots(cm, …)`: the channel manager of this box behind its unix socket
    (`ChannelManager.serve_local` in that process).  One transport per process: connecting installs the client table of the
    slots runtime's seam, like `grpc_api.remote_cm.connect` does for the gRPC transport."""

    def __init__(self, socket_path: str):
        h = _H()
        _check(lib.lzyb_slots_connect_cm_local(_b(socket_path), C.byref(h)))
        self.h, self.socket_path = h.value, socket_path

    def close(self) -> None:
        if self.h:
            lib.lzyb_slots_disconnect_cm_local()
            self.h = 0


def pump_start(cm: ChannelManager) -> None:
    _check(lib.lzyb_cm_pump_start(cm.h))


def pump_stop(cm: ChannelManager) -> None:
    _check(lib.lzyb_cm_pump_stop(cm.h))
