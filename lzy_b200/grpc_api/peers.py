"""Consumer side of peers in other worker processes: the slots runtime asks (include/lzy_b200_host.h:
``lzyb_slots_set_peer_resolver``) for the descriptor of a peer it cannot find locally; the answer comes from that
peer's worker over ``LzybPeerApi.Describe``.  ``install()`` once per process."""
from __future__ import annotations

import ctypes as C
import threading
from typing import Callable, Dict, Optional

import grpc

from .. import _host as H
from .. import _native as N
from .client import PeerApiClient
from .protos import DescribeRequest
from .pump import _target_of
from .remote_cm import _CODES

_RESOLVER = C.CFUNCTYPE(C.c_int, C.c_char_p, C.c_char_p, C.c_void_p)
_lock = threading.Lock()
_clients: Dict[str, PeerApiClient] = {}
_installed = None
_to_target: Callable[[str], str] = _target_of
resolved = 0  # how many descriptors were fetched (tests, metrics)


def _resolve(url: bytes, peer_id: bytes, out) -> int:
    global resolved
    try:
        target = _to_target((url or b"").decode())
        with _lock:
            client = _clients.get(target)
            if client is None:
                client = _clients[target] = PeerApiClient(target)
        raw = client.call("Describe", DescribeRequest(peer_id=(peer_id or b"").decode()), timeout=10.0).descriptor
    except grpc.RpcError as e:
        N.lib.lzyb_set_last_error((e.details() or str(e.code())).encode())
        return _CODES.get(e.code(), N.E_PEER)
    except Exception as e:  # noqa: BLE001 — an exception escaping a ctypes callback would read as rc 0
        N.lib.lzyb_set_last_error(f"peer resolver failed: {type(e).__name__}: {e}".encode()[:500])
        return N.E_PEER
    if len(raw) != C.sizeof(N.IpcDesc):
        N.lib.lzyb_set_last_error(b"peer sent a malformed descriptor")
        return N.E_PEER
    C.memmove(out, raw, len(raw))
    resolved += 1
    return N.OK


def install(url_to_target: Optional[Callable[[str], str]] = None) -> None:
    global _installed, _to_target
    with _lock:
        if url_to_target is not None:
            _to_target = url_to_target
        if _installed is None:
            _installed = _RESOLVER(_resolve)
            H._check(H.lib.lzyb_slots_set_peer_resolver(C.cast(_installed, C.c_void_p)))


def uninstall() -> None:
    global _installed
    with _lock:
        H.lib.lzyb_slots_set_peer_resolver(None)
        _installed = None
        for c in _clients.values():
            c.close()
        _clients.clear()
