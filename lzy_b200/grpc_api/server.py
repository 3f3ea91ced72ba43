"""gRPC servers for the kept surface, backed by the C++ host logic (``_host.Slots`` /
``_host.ChannelManager``).  Status codes are the ones the reference returns at the same places
(lzy/slots/.../SlotsService.java:36,44,55,63; channel-manager/.../services/SlotsService.java); the
``Idempotency-Key`` header (util/util-grpc/.../GrpcHeaders.java:17) is honoured on mutating calls."""
from __future__ import annotations

from concurrent import futures
from typing import Callable, Dict, Iterable, Optional

import grpc

from .. import _host as H
from .. import _native as N
from .protos import SERVICES, DescribeResponse, LongRunningOperation, messages, method_types

READ_CHUNK = 2 << 20  # OutputSlot.java:241
WAIT_HEADER = "lzyb-wait-ms"  # LongRunningService.Get: hold the answer until the operation is done (at most that long)
_STATUS = {N.E_NOT_FOUND: grpc.StatusCode.NOT_FOUND, N.E_PRECONDITION: grpc.StatusCode.FAILED_PRECONDITION,
           N.E_UNAVAILABLE: grpc.StatusCode.UNAVAILABLE, N.E_INTERNAL: grpc.StatusCode.INTERNAL,
           N.E_UNIMPLEMENTED: grpc.StatusCode.UNIMPLEMENTED, N.E_INVALID: grpc.StatusCode.INVALID_ARGUMENT,
           N.E_EXISTS: grpc.StatusCode.ALREADY_EXISTS}


def _abort(context, err: N.LzybError):
    context.abort(_STATUS.get(err.code, grpc.StatusCode.INTERNAL), err.message)


def _idk(context) -> Optional[str]:
    for k, v in context.invocation_metadata() or ():
        if k.lower() == "idempotency-key":
            return v
    return None


def _peer_to_proto(peer: H.Peer):
    p = messages.PeerDescription(peer_id=peer.peer_id)
    if peer.is_storage:
        p.storage_peer.storage_uri = peer.url
    else:
        p.slot_peer.peer_url = peer.url
    return p


def _peer_from_proto(p) -> H.Peer:
    if p.WhichOneof("peer") == "storage_peer":
        return H.Peer(p.peer_id, H.PEER_STORAGE, p.storage_peer.storage_uri)
    return H.Peer(p.peer_id, H.PEER_SLOT, p.slot_peer.peer_url)


def _status_to_proto(st: dict):
    cs = messages.ChannelStatus(channel_id=st["channel_id"], execution_id=st["execution_id"])
    cs.scheme.dataFormat = st["scheme"]["data_format"]
    cs.scheme.schemeFormat = st["scheme"]["scheme_format"]
    cs.scheme.schemeContent = st["scheme"]["scheme_content"]
    for key, dst in (("producers", cs.producers), ("consumers", cs.consumers)):
        for p in st[key]:
            dst.append(_peer_to_proto(H.Peer(p["peer_id"], H.PEER_STORAGE if p["kind"] == "storage" else H.PEER_SLOT, p["url"])))
    return cs


class SlotsApiService:
    """`slots` is a `_host.Slots`, or a `lzy_b200.worker.Worker` whose slots server exists only after
    WorkerApi.Init — until then both calls answer UNAVAILABLE (WorkerTests.java:261-268)."""
    SERVICE = "ai.lzy.v1.slots.LzySlotsApi"

    def __init__(self, slots):
        self._source = slots

    @property
    def slots(self) -> H.Slots:
        return self._source if isinstance(self._source, H.Slots) else self._source.slots

    def _available(self, context) -> None:
        if self.slots is None:
            context.abort(grpc.StatusCode.UNAVAILABLE, "worker is not initialized")

    def StartTransfer(self, request, context):
        self._available(context)
        try:
            self.slots.start_transfer(request.slot_id, _peer_from_proto(request.peer), request.transfer_id)
        except N.LzybError as e:
            _abort(context, e)
        return messages.StartTransferResponse()

    def Read(self, request, context) -> Iterable:
        self._available(context)
        offset = int(request.offset)
        try:
            while True:
                chunk = self.slots.read(request.peer_id, offset, READ_CHUNK)
                if chunk is None:
                    break
                offset += len(chunk)
                yield messages.ReadDataChunk(chunk=chunk)
        except N.LzybError as e:
            _abort(context, e)
        yield messages.ReadDataChunk(control=0)  # Control.EOS


class ChannelManagerService:
    SERVICE = "ai.lzy.v1.channel.LzyChannelManager"
    PRIVATE_SERVICE = "ai.lzy.v1.channel.LzyChannelManagerPrivate"

    def __init__(self, cm: H.ChannelManager):
        self.cm = cm

    # ---- public ----
    def Bind(self, request, context):
        role = {1: H.ROLE_CONSUMER, 2: H.ROLE_PRODUCER}.get(int(request.role), 0)
        try:
            peer, tid = self.cm.bind(request.channel_id, role, request.peer_id, request.peer_url, request.execution_id, _idk(context))
        except N.LzybError as e:
            _abort(context, e)
        resp = messages.BindResponse()
        if peer is not None:
            resp.peer.CopyFrom(_peer_to_proto(peer))
            resp.transfer_id = tid
        return resp

    def Unbind(self, request, context):
        try:
            self.cm.unbind(request.channel_id, request.peer_id)
        except N.LzybError as e:
            _abort(context, e)
        return messages.UnbindResponse()

    def TransferCompleted(self, request, context):
        try:
            self.cm.transfer_completed(request.channel_id, request.transfer_id, _idk(context))
        except N.LzybError as e:
            _abort(context, e)
        return messages.TransferCompletedResponse()

    def TransferFailed(self, request, context):
        try:
            peer, tid = self.cm.transfer_failed(request.channel_id, request.transfer_id, request.description, _idk(context))
        except N.LzybError as e:
            _abort(context, e)
        resp = messages.TransferFailedResponse(new_transfer_id=tid)
        resp.new_peer.CopyFrom(_peer_to_proto(peer))
        return resp

    def GetChannelsStatus(self, request, context):
        resp = messages.GetChannelsStatusResponse()
        for st in self.cm.channels_status(request.execution_id, list(request.channel_ids)):
            resp.channels.append(_status_to_proto(st))
        return resp

    # ---- private ----
    def GetOrCreate(self, request, context):
        which = request.WhichOneof("initial_storage_peer")
        kw = {}
        if which == "producer":
            kw["producer_uri"] = request.producer.storage_uri
        elif which == "consumer":
            kw["consumer_uri"] = request.consumer.storage_uri
        try:
            cid = self.cm.get_or_create(request.user_id, request.workflow_name, request.execution_id,
                                        data_format=request.scheme.dataFormat, scheme_format=request.scheme.schemeFormat,
                                        scheme_content=request.scheme.schemeContent, **kw)
        except N.LzybError as e:
            _abort(context, e)
        return messages.GetOrCreateResponse(channel_id=cid)

    def Destroy(self, request, context):
        self.cm.destroy(request.channel_id, request.reason)
        return messages.DestroyResponse()

    def DestroyAll(self, request, context):
        self.cm.destroy_all(request.execution_id, request.reason)
        return messages.DestroyAllResponse()

    def Status(self, request, context):
        sts = self.cm.channels_status("", [request.channel_id])
        if not sts:
            context.abort(grpc.StatusCode.NOT_FOUND, f"Channel {request.channel_id} not found")
        resp = messages.StatusResponse()
        resp.channel.CopyFrom(_status_to_proto(sts[0]))
        return resp

    def List(self, request, context):
        resp = messages.ListResponse()
        for st in self.cm.channels_status(request.execution_id):
            resp.statuses.append(_status_to_proto(st))
        return resp


class PeerApiService:
    """ai.lzy.b200.v1.LzybPeerApi — the additive RPC a consumer in another worker process uses to map a producer's
    frame: Describe(peer_id) → lzyb_ipc_desc_t (same NOT_FOUND / UNAVAILABLE behaviour as LzySlotsApi.Read)."""
    SERVICE = "ai.lzy.b200.v1.LzybPeerApi"

    def __init__(self, slots):
        self._source = slots

    def Describe(self, request, context):
        slots = self._source if isinstance(self._source, H.Slots) else self._source.slots
        if slots is None:
            context.abort(grpc.StatusCode.UNAVAILABLE, "worker is not initialized")
        try:
            return DescribeResponse(descriptor=slots.describe(request.peer_id))
        except N.LzybError as e:
            _abort(context, e)


def _wait_budget(context) -> float:
    """seconds the caller allows this call to wait for the operation (`lzyb-wait-ms`, never past the call's own deadline);
    0 without the header — then the answer comes at once, as from the reference's services"""
    wait_ms = 0.0
    for k, v in context.invocation_metadata() or ():
        if k.lower() == WAIT_HEADER:
            try:
                wait_ms = max(0.0, min(float(v), 600_000.0))
            except ValueError:
                context.abort(grpc.StatusCode.INVALID_ARGUMENT, f"{WAIT_HEADER}: not a number: {v!r}")
    if wait_ms <= 0:
        return 0.0
    left = context.time_remaining()
    return wait_ms * 1e-3 if left is None else max(0.0, min(wait_ms * 1e-3, left - 0.05))


def _operation_to_proto(op):
    """worker.Operation → ai.lzy.v1.longrunning.Operation; the result is ``response = Any(ExecuteResponse)``
    as declared by the ``(ai.lzy.v1.operation)`` option of WorkerApi.Execute (worker-service.proto:17-22)"""
    p = LongRunningOperation(id=op.id, description=op.description, created_by=op.created_by, done=op.done)
    p.created_at.FromNanoseconds(int(op.created_at * 1e9))
    p.modified_at.FromNanoseconds(int(op.modified_at * 1e9))
    if op.error is not None:
        p.error.code, p.error.message = op.error
    elif op.response is not None:
        p.response.Pack(messages.ExecuteResponse(rc=op.response.rc, description=op.response.description))
    return p


def _worker_abort(context, err):
    context.abort(getattr(grpc.StatusCode, err.code, grpc.StatusCode.INTERNAL), err.message)


class WorkerApiService:
    """ai.lzy.v1.worker.WorkerApi over `lzy_b200.worker.Worker` (WorkerApiImpl.java:85-141,232-283)."""
    SERVICE = "ai.lzy.v1.worker.WorkerApi"

    def __init__(self, worker):
        self.worker = worker

    def Init(self, request, context):
        from ..worker import WorkerStatusError
        try:
            self.worker.init(request.user_id, request.workflow_name, request.worker_subject_name, request.worker_private_key)
        except WorkerStatusError as e:
            _worker_abort(context, e)
        return messages.InitResponse()

    def Execute(self, request, context):
        from ..worker import SlotSpec, TaskSpec, WorkerStatusError
        op = request.task_desc.operation
        task = TaskSpec(task_id=request.task_id, execution_id=request.execution_id, workflow_name=request.workflow_name,
                        user_id=request.user_id, command=op.command, name=op.name, env=op.env,
                        slots=[SlotSpec(s.name, int(s.direction)) for s in op.slots],
                        slot_assignments={a.slot_name: a.channel_id for a in request.task_desc.slot_assignments})
        try:
            op_state = self.worker.execute(task, _idk(context))
            wait_s = _wait_budget(context)
            if wait_s > 0 and not op_state.done:
                # the caller asked (`lzyb-wait-ms`) to be answered when the task is done: the long-running operation comes back
                # completed — `done` with its response, which every client of the operation pattern must accept anyway — and
                # the driver saves the Get round trip (LongRunningService.Get honours the same header)
                op_state = self.worker.wait_operation(op_state.id, wait_s)
            return _operation_to_proto(op_state)
        except WorkerStatusError as e:
            _worker_abort(context, e)


class LongRunningService:
    """ai.lzy.v1.longrunning.LongRunningService.Get for the worker's operations (what the graph executor polls)."""
    SERVICE = "ai.lzy.v1.longrunning.LongRunningService"

    def __init__(self, worker):
        self.worker = worker

    def Get(self, request, context):
        """Answers at once, as the reference's operation service does — unless the caller sent ``lzyb-wait-ms: <n>``: then an
        operation that is still running is answered when it completes, or as it stands after n ms (never past the call's own
        deadline).  A driver on the same box gets the completion the moment it happens without polling the worker's
        interpreter every fraction of a millisecond (tools/worker_chain_cpu.py: each poll costs the running task ≈ 0.3 ms)."""
        from ..worker import WorkerStatusError
        wait_s = _wait_budget(context)
        try:
            if wait_s > 0:
                return _operation_to_proto(self.worker.wait_operation(request.operation_id, wait_s))
            return _operation_to_proto(self.worker.get_operation(request.operation_id))
        except WorkerStatusError as e:
            _worker_abort(context, e)


def _generic_handler(service: str, impl) -> grpc.GenericRpcHandler:
    handlers: Dict[str, grpc.RpcMethodHandler] = {}
    for method in SERVICES[service]:
        req_cls, resp_cls, streaming = method_types(service, method)
        fn: Callable = getattr(impl, method)
        mk = grpc.unary_stream_rpc_method_handler if streaming else grpc.unary_unary_rpc_method_handler
        handlers[method] = mk(fn, request_deserializer=req_cls.FromString, response_serializer=resp_cls.SerializeToString)
    return grpc.method_handlers_generic_handler(service, handlers)


def serve(address: str = "127.0.0.1:0", slots: Optional[H.Slots] = None, cm: Optional[H.ChannelManager] = None,
          max_workers: int = 8, worker=None):
    """start a server with whichever services were given; returns (server, bound port)"""
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers),
                         options=[("grpc.max_send_message_length", 8 << 20), ("grpc.max_receive_message_length", 8 << 20)])
    if slots is not None:
        server.add_generic_rpc_handlers((_generic_handler(SlotsApiService.SERVICE, SlotsApiService(slots)),
                                         _generic_handler(PeerApiService.SERVICE, PeerApiService(slots))))
    if cm is not None:
        svc = ChannelManagerService(cm)
        server.add_generic_rpc_handlers((_generic_handler(svc.SERVICE, svc), _generic_handler(svc.PRIVATE_SERVICE, svc)))
    if worker is not None:
        if slots is None:
            server.add_generic_rpc_handlers((_generic_handler(SlotsApiService.SERVICE, SlotsApiService(worker)),
                                             _generic_handler(PeerApiService.SERVICE, PeerApiService(worker))))
        server.add_generic_rpc_handlers((_generic_handler(WorkerApiService.SERVICE, WorkerApiService(worker)),
                                         _generic_handler(LongRunningService.SERVICE, LongRunningService(worker))))
    port = server.add_insecure_port(address)
    server.start()
    return server, port
