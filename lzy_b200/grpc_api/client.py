"""Thin blocking clients for the kept gRPC surface (what the reference's generated stubs do)."""
from __future__ import annotations

from typing import Optional

import grpc

from .protos import SERVICES, method_types


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

    def call(self, method: str, request, idempotency_key: Optional[str] = None, timeout: float = 30.0):
        md = (("idempotency-key", idempotency_key),) if idempotency_key else None
        return getattr(self, "_" + method)(request, metadata=md, timeout=timeout)

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


class LongRunningClient(_Client):
    SERVICE = "ai.lzy.v1.longrunning.LongRunningService"


class PeerApiClient(_Client):
    SERVICE = "ai.lzy.b200.v1.LzybPeerApi"
