"""Thin blocking clients for the kept gRPC surface (what the reference's generated stubs do)."""
from __future__ import annotations

import os
import sys
import time
from typing import Optional

import grpc

from .protos import SERVICES, method_types

_TRACE = bool(os.environ.get("LZYB_TRACE"))


class _Client:
    SERVICE = ""

    def __init__(self, target: str):
        self.target = target
        self.channel = grpc.insecure_channel(target, options=[("grpc.max_receive_message_length", 8 << 20)])
        for method in SERVICES[self.SERVICE]:
            req_cls, resp_cls, streaming = method_types(self.SERVICE, method)
            mk = self.channel.unary_stream if streaming else self.channel.unary_unary
            setattr(self, "_" + method, mk(f"/{self.SERVICE}/{method}", request_serializer=req_cls.SerializeToString,
                                           response_deserializer=resp_cls.FromString))

    def call(self, method: str, request, idempotency_key: Optional[str] = None, timeout: float = 30.0,
             metadata: Optional[tuple] = None):
        md = (("idempotency-key", idempotency_key),) if idempotency_key else None
        if metadata:
            md = (md or ()) + tuple(metadata)
        if not _TRACE:
            return getattr(self, "_" + method)(request, metadata=md, timeout=timeout)
        t0 = time.perf_counter()  # LZYB_TRACE=1: one stderr line per outgoing call (next to the worker's phase lines)
        try:
            return getattr(self, "_" + method)(request, metadata=md, timeout=timeout)
        finally:
            print(f"LZYB_TRACE rpc={method} ms={1e3 * (time.perf_counter() - t0):.3f} t={t0:.6f}", file=sys.stderr, flush=True)

    def close(self):
        self.channel.close()


class SlotsApiClient(_Client):
    SERVICE = "ai.lzy.v1.slots.LzySlotsApi"

    def read_all(self, peer_id: str, offset: int = 0) -> bytes:
        """a legacy consumer: concatenates the chunks of one Read stream up to EOS"""
        from .protos import messages
        out = bytearray()
        saw_eos = False
        for chunk in self.call("Read", messages.ReadDataRequest(peer_id=peer_id, offset=offset)):
            if chunk.WhichOneof("kind") == "control":
                saw_eos = True
                break
            out += chunk.chunk
        if not saw_eos:
            raise RuntimeError("stream ended without EOS")
        return bytes(out)


class ChannelManagerClient(_Client):
    SERVICE = "ai.lzy.v1.channel.LzyChannelManager"


class ChannelManagerPrivateClient(_Client):
    SERVICE = "ai.lzy.v1.channel.LzyChannelManagerPrivate"


class WorkerApiClient(_Client):
    SERVICE = "ai.lzy.v1.worker.WorkerApi"

    def execute_and_wait(self, request, timeout: float = 300.0, idempotency_key: Optional[str] = None):
        """``Execute`` with the ``lzyb-wait-ms`` header: this repository's workers answer with the COMPLETED operation (one
        round trip per task); a worker that ignores the header answers at once with a running one — then ask
        ``LongRunningClient.wait``"""
        return self.call("Execute", request, idempotency_key=idempotency_key, timeout=timeout + 10.0,
                         metadata=(("lzyb-wait-ms", str(int(timeout * 1e3))),))


class LongRunningClient(_Client):
    SERVICE = "ai.lzy.v1.longrunning.LongRunningService"

    def wait(self, operation_id: str, timeout: float = 300.0, slice_ms: int = 1000):
        """the operation once it is done: ``Get`` with the ``lzyb-wait-ms`` header, which this repository's workers answer
        when the operation completes (a server that ignores the header answers at once: then this is a poll with a
        short sleep, as against the reference's operation service)"""
        from .protos import GetOperationRequest
        deadline = time.monotonic() + timeout
        req = GetOperationRequest(operation_id=operation_id)
        while True:
            t0 = time.monotonic()
            op = self.call("Get", req, timeout=slice_ms * 1e-3 + 10.0, metadata=(("lzyb-wait-ms", str(slice_ms)),))
            if op.done:
                return op
            if time.monotonic() > deadline:
                raise TimeoutError(f"operation {operation_id} still running after {timeout} s")
            if time.monotonic() - t0 < slice_ms * 5e-4:  # back long before the slice was over: the header was not honoured
                time.sleep(2e-3)


class PeerApiClient(_Client):
    SERVICE = "ai.lzy.b200.v1.LzybPeerApi"
