"""A scripted channel manager for the slot state machines — the role ``ChannelManagerMock`` plays in
lzy/slots/src/test/java/ai/lzy/slots/SlotsTest.java: every call the slots make (Bind / Unbind / TransferCompleted /
TransferFailed) is recorded, and a test may take a *handle* on a call before it happens, wait until the slot really
makes it, look at the request, and only then answer it (with a peer, an empty response or a gRPC failure).  Calls
nobody scripted get the default empty answer.  Installed at the client seam of the slots runtime
(include/lzy_b200_host.h: ``lzyb_slots_set_cm_client``); test infrastructure only."""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, List, Optional, Tuple

from lzy_b200 import _host as H
from lzy_b200 import _native as N
from lzy_b200.grpc_api import remote_cm as R

COOKIE = 0x5C12197ED


class Handle:
    """one expected call: ``get()`` waits for the slot to make it and returns the request; ``complete(...)`` /
    ``complete_exceptionally(...)`` lets it return"""

    def __init__(self):
        self._arrived, self._answered = threading.Event(), threading.Event()
        self.request: Optional[dict] = None
        self._peer: Optional[H.Peer] = None
        self._transfer_id = ""
        self._rc = N.OK
        self._message = ""

    def get(self, timeout: float = 10.0) -> dict:
        assert self._arrived.wait(timeout), "the slot never made this call"
        return self.request

    def complete(self, peer: Optional[H.Peer] = None, transfer_id: str = "") -> "Handle":
        self._peer, self._transfer_id = peer, transfer_id
        self._answered.set()
        return self

    def complete_exceptionally(self, message: str, rc: int = N.E_INTERNAL) -> "Handle":
        self._rc, self._message = rc, message
        self._answered.set()
        return self

    # called on the slot's thread
    def _serve(self, request: dict, out=None) -> int:
        self.request = request
        self._arrived.set()
        assert self._answered.wait(30.0), f"the test never answered {request}"
        if self._rc != N.OK:
            N.lib.lzyb_set_last_error(self._message.encode())
            return self._rc
        if out is not None:
            _fill(out, self._peer, self._transfer_id)
        return N.OK


def _fill(out, peer: Optional[H.Peer], transfer_id: str) -> None:
    C.memset(out, 0, C.sizeof(H.BindResult))
    if peer is None:
        return
    r = out.contents
    r.has_peer = 1
    r.peer.peer_id, r.peer.kind, r.peer.url = peer.peer_id.encode(), peer.kind, peer.url.encode()
    r.transfer_id = transfer_id.encode()


class ScriptedChannelManager:
    def __init__(self):
        self.h = COOKIE
        self.calls: List[Tuple[str, dict]] = []
        self._lock = threading.Lock()
        self._scripts: Dict[Tuple[str, str], List[Handle]] = {}
        self._ops = R._ClientOps(R._BIND(self._bind), R._UNBIND(self._unbind), R._COMPLETED(self._completed),
                                 R._FAILED(self._failed))
        H._check(H.lib.lzyb_slots_set_cm_client(C.byref(self._ops)))

    def close(self) -> None:
        H.lib.lzyb_slots_set_cm_client(None)

    # --- scripting ---------------------------------------------------------------------------------------------
    def _expect(self, kind: str, key: str) -> Handle:
        h = Handle()
        with self._lock:
            self._scripts.setdefault((kind, key), []).append(h)
        return h

    def on_bind(self, peer_id: str) -> Handle:
        return self._expect("bind", peer_id)

    def on_unbind(self, peer_id: str) -> Handle:
        return self._expect("unbind", peer_id)

    def on_transfer_completed(self, transfer_id: str) -> Handle:
        return self._expect("completed", transfer_id)

    def on_transfer_failed(self, transfer_id: str) -> Handle:
        return self._expect("failed", transfer_id)

    def made(self, kind: str) -> List[dict]:
        with self._lock:
            return [req for k, req in self.calls if k == kind]

    # --- the client table --------------------------------------------------------------------------------------
    def _dispatch(self, kind: str, key: str, request: dict, out=None) -> int:
        try:
            with self._lock:
                self.calls.append((kind, request))
                queue = self._scripts.get((kind, key))
                handle = queue.pop(0) if queue else None
            if handle is None:
                if out is not None:
                    _fill(out, None, "")
                return N.OK
            return handle._serve(request, out)
        except Exception as e:  # noqa: BLE001 — an exception escaping a ctypes callback would read as rc 0
            N.lib.lzyb_set_last_error(f"scripted channel manager: {type(e).__name__}: {e}".encode()[:500])
            return N.E_INTERNAL

    def _bind(self, cookie, idk, channel_id, peer_id, execution_id, peer_url, role, out) -> int:
        return self._dispatch("bind", R._s(peer_id), dict(
            channel_id=R._s(channel_id), peer_id=R._s(peer_id), execution_id=R._s(execution_id), peer_url=R._s(peer_url),
            role=role, idempotency_key=R._s(idk), cookie=cookie), out)

    def _unbind(self, cookie, channel_id, peer_id) -> int:
        return self._dispatch("unbind", R._s(peer_id), dict(channel_id=R._s(channel_id), peer_id=R._s(peer_id)))

    def _completed(self, cookie, idk, channel_id, transfer_id) -> int:
        return self._dispatch("completed", R._s(transfer_id), dict(
            channel_id=R._s(channel_id), transfer_id=R._s(transfer_id), idempotency_key=R._s(idk)))

    def _failed(self, cookie, idk, channel_id, transfer_id, description, out) -> int:
        return self._dispatch("failed", R._s(transfer_id), dict(
            channel_id=R._s(channel_id), transfer_id=R._s(transfer_id), description=R._s(description),
            idempotency_key=R._s(idk)), out)
