"""A worker whose channel manager is another process: the four calls the slot state machines make into the matcher
(``LzyChannelManager.Bind / Unbind / TransferCompleted / TransferFailed``, channel-manager.proto:13-58) become RPCs.
`connect(target)` installs the client table at the slots runtime's seam (include/lzy_b200_host.h:
``lzyb_slots_set_cm_client``) and returns an object to hand to ``_host.Slots`` in place of a local
``ChannelManager``; status codes come back as the LZYB_E_* images of the gRPC codes, so the state machines react to
them exactly as to the in-process matcher (``UNAVAILABLE`` on Unbind while transfers are live,
``FAILED_PRECONDITION`` for a non-active transfer, ``INTERNAL`` when no producer is left …)."""
from __future__ import annotations

import ctypes as C
import itertools
import threading
from typing import Dict, Optional

import grpc

from .. import _host as H
from .. import _native as N
from .client import ChannelManagerClient
from .protos import messages

_CODES = {grpc.StatusCode.NOT_FOUND: N.E_NOT_FOUND, grpc.StatusCode.FAILED_PRECONDITION: N.E_PRECONDITION,
          grpc.StatusCode.UNAVAILABLE: N.E_UNAVAILABLE, grpc.StatusCode.INTERNAL: N.E_INTERNAL,
          grpc.StatusCode.UNIMPLEMENTED: N.E_UNIMPLEMENTED, grpc.StatusCode.INVALID_ARGUMENT: N.E_INVALID,
          grpc.StatusCode.ALREADY_EXISTS: N.E_EXISTS}
_ROLE_TO_WIRE = {H.ROLE_CONSUMER: 1, H.ROLE_PRODUCER: 2}  # LCMS.BindRequest.Role

_BIND = C.CFUNCTYPE(C.c_int, C.c_uint64, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(H.BindResult))
_UNBIND = C.CFUNCTYPE(C.c_int, C.c_uint64, C.c_char_p, C.c_char_p)
_COMPLETED = C.CFUNCTYPE(C.c_int, C.c_uint64, C.c_char_p, C.c_char_p, C.c_char_p)
_FAILED = C.CFUNCTYPE(C.c_int, C.c_uint64, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(H.BindResult))


class _ClientOps(C.Structure):
    _fields_ = [("bind", _BIND), ("unbind", _UNBIND), ("transfer_completed", _COMPLETED), ("transfer_failed", _FAILED)]


class RemoteChannelManager:
    """stands where a local ``_host.ChannelManager`` would: ``.h`` is the cookie the client table routes by"""

    def __init__(self, cookie: int, client: ChannelManagerClient, target: str):
        self.h, self.client, self.target = cookie, client, target

    def close(self) -> None:
        with _lock:
            _remotes.pop(self.h, None)
        self.client.close()


_lock = threading.Lock()
_remotes: Dict[int, RemoteChannelManager] = {}
_cookies = itertools.count(0x5EED0001)
_installed: Optional[_ClientOps] = None


def _s(b: Optional[bytes]) -> str:
    return (b or b"").decode()


def _fail(e: grpc.RpcError) -> int:
    N.lib.lzyb_set_last_error((e.details() or str(e.code())).encode())
    return _CODES.get(e.code(), N.E_INTERNAL)


def _fill(out, peer, transfer_id: str) -> None:
    C.memset(out, 0, C.sizeof(H.BindResult))
    r = out.contents
    which = peer.WhichOneof("peer") if peer is not None else None
    if which is None:
        return
    r.has_peer = 1
    r.peer.peer_id = peer.peer_id.encode()
    if which == "storage_peer":
        r.peer.kind, r.peer.url = H.PEER_STORAGE, peer.storage_peer.storage_uri.encode()
    else:
        r.peer.kind, r.peer.url = H.PEER_SLOT, peer.slot_peer.peer_url.encode()
    r.transfer_id = transfer_id.encode()


def _guarded(fn):
    """an exception escaping a ctypes callback would be printed and read as rc 0: turn it into INTERNAL"""
    def wrapper(*args):
        try:
            return fn(*args)
        except Exception as e:  # noqa: BLE001
            N.lib.lzyb_set_last_error(f"channel manager client failed: {type(e).__name__}: {e}".encode()[:500])
            return N.E_INTERNAL
    return wrapper


@_guarded
def _bind(cookie, idk, channel_id, peer_id, execution_id, peer_url, role, out) -> int:
    try:
        resp = _remotes[cookie].client.call("Bind", messages.BindRequest(
            execution_id=_s(execution_id), channel_id=_s(channel_id), peer_id=_s(peer_id), peer_url=_s(peer_url),
            role=_ROLE_TO_WIRE.get(role, 0)), idempotency_key=_s(idk) or None)
    except grpc.RpcError as e:
        return _fail(e)
    _fill(out, resp.peer if resp.HasField("peer") else None, resp.transfer_id)
    return N.OK


@_guarded
def _unbind(cookie, channel_id, peer_id) -> int:
    try:
        _remotes[cookie].client.call("Unbind", messages.UnbindRequest(peer_id=_s(peer_id), channel_id=_s(channel_id)))
    except grpc.RpcError as e:
        return _fail(e)
    return N.OK


@_guarded
def _completed(cookie, idk, channel_id, transfer_id) -> int:
    try:
        _remotes[cookie].client.call("TransferCompleted", messages.TransferCompletedRequest(
            transfer_id=_s(transfer_id), channel_id=_s(channel_id)), idempotency_key=_s(idk) or None)
    except grpc.RpcError as e:
        return _fail(e)
    return N.OK


@_guarded
def _failed(cookie, idk, channel_id, transfer_id, description, out) -> int:
    try:
        resp = _remotes[cookie].client.call("TransferFailed", messages.TransferFailedRequest(
            transfer_id=_s(transfer_id), channel_id=_s(channel_id), description=_s(description)), idempotency_key=_s(idk) or None)
    except grpc.RpcError as e:
        return _fail(e)
    _fill(out, resp.new_peer if resp.HasField("new_peer") else None, resp.new_transfer_id)
    return N.OK


def connect(target: str) -> RemoteChannelManager:
    """route this process's slot state machines to the channel manager at `target`"""
    global _installed
    with _lock:
        if _installed is None:
            _installed = _ClientOps(_BIND(_bind), _UNBIND(_unbind), _COMPLETED(_completed), _FAILED(_failed))
            H._check(H.lib.lzyb_slots_set_cm_client(C.byref(_installed)))
        remote = RemoteChannelManager(next(_cookies), ChannelManagerClient(target), target)
        _remotes[remote.h] = remote
        return remote


def disconnect_all() -> None:
    """back to the in-process matcher"""
    global _installed
    with _lock:
        H.lib.lzyb_slots_set_cm_client(None)
        _installed = None
        for r in list(_remotes.values()):
            r.client.close()
        _remotes.clear()
