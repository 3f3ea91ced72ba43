"""The channel manager's deferred actions delivered over the wire — ``ActionScheduler.startTransfer``
(lzy/channel-manager/src/main/java/ai/lzy/channelmanager/ActionScheduler.java:61-106) for a channel manager that
runs as its own service: every ``START_TRANSFER`` action becomes ``LzySlotsApi.StartTransfer(slot_id = consumer,
peer = producer, transfer_id)`` sent to the consumer's worker, every ``ABORT_WORKFLOW`` action is handed to a callback
(the reference calls lzy-service, LzyServiceClient.java:37-55).

Two deliberate differences, both already made by the in-process pump (csrc/host/slots_runtime.cpp: pump_loop):
the request is sent to the *consumer's* URL (the reference dials the producer's, ActionScheduler.java:64, which
only works when both slots share one slots server), and the transfer is marked ACTIVE before the request goes out —
a consumer on NVLink can finish and report TransferCompleted before the reference's "mark active afterwards"
(:79-91) would have run; a failed delivery then fails the channel as in the reference (:92-104).
"""
from __future__ import annotations

import logging
import threading
from typing import Callable, Dict, Optional

import grpc

from .. import _host as H
from .client import SlotsApiClient
from .protos import messages

_LOG = logging.getLogger(__name__)


def _target_of(url: str) -> str:
    """worker URL → gRPC target: ``host:port`` as is, scheme prefixes such as ``cuda+ipc://host:port?gpu=1`` stripped"""
    rest = url.split("://", 1)[-1]
    return rest.split("?", 1)[0].split("/", 1)[0]


class RemoteActionPump:
    def __init__(self, cm: H.ChannelManager, on_abort: Optional[Callable[[dict], None]] = None,
                 resolve: Callable[[str], str] = _target_of, poll_ms: int = 20):
        self.cm, self.on_abort, self.resolve, self.poll_ms = cm, on_abort, resolve, poll_ms
        self._clients: Dict[str, SlotsApiClient] = {}
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self.delivered = 0
        self.failed = 0

    def start(self) -> "RemoteActionPump":
        self.cm.restore_actions()  # PENDING transfers of a previous incarnation (ActionScheduler.java:47-59)
        self._thread = threading.Thread(target=self._loop, name="lzyb-cm-pump", daemon=True)
        self._thread.start()
        return self

    def stop(self) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=10)
        for c in self._clients.values():
            c.close()
        self._clients.clear()

    def _client(self, url: str) -> SlotsApiClient:
        target = self.resolve(url)
        c = self._clients.get(target)
        if c is None:
            c = self._clients[target] = SlotsApiClient(target)
        return c

    def _loop(self) -> None:
        while not self._stop.is_set():
            action = self.cm.poll_action(self.poll_ms)
            if action is None:
                continue
            if action["kind"] == H.ACT_START_TRANSFER:
                self._start_transfer(action)
            elif action["kind"] == H.ACT_ABORT_WORKFLOW and self.on_abort is not None:
                try:
                    self.on_abort(action)
                except Exception:  # noqa: BLE001 — a failing callback must not stop the pump
                    _LOG.exception("abort callback failed for execution %s", action["execution_id"])

    def _start_transfer(self, a: dict) -> None:
        self.cm.action_result(a["channel_id"], a["transfer_id"], True)
        req = messages.StartTransferRequest(transfer_id=a["transfer_id"], slot_id=a["consumer_id"])
        peer: H.Peer = a["producer"]
        req.peer.peer_id = peer.peer_id
        if peer.is_storage:
            req.peer.storage_peer.storage_uri = peer.url
        else:
            req.peer.slot_peer.peer_url = peer.url
        try:
            self._client(a["consumer_url"]).call("StartTransfer", req, idempotency_key=a["idempotency_key"] or None, timeout=10.0)
            self.delivered += 1
        except grpc.RpcError as e:
            self.failed += 1
            _LOG.error("Cannot connect to %s for transfer %s: %s", a["consumer_url"], a["transfer_id"], e)
            self.cm.action_result(a["channel_id"], a["transfer_id"], False, "Cannot connect.")
