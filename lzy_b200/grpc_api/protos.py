"""Protobuf messages of the kept wire contract, declared programmatically (this image has no
``protoc``/``grpc_tools``).  Field names and numbers restate, message by message:

  ai/lzy/v1/common/data-scheme.proto:8-14      DataScheme
  ai/lzy/v1/common/storage.proto:10-18         S3Credentials, AzureBlobStorageCredentials
  ai/lzy/v1/common/channel.proto:12-40         ChannelStatus, PeerDescription{SlotPeer, StoragePeer}
  ai/lzy/v1/slots/slots-api.proto:23-47        StartTransferRequest/Response, ReadDataRequest, ReadDataChunk
  ai/lzy/v1/channel/channel-manager.proto:27-95            Bind*, Unbind*, GetChannelsStatus*, TransferCompleted*, TransferFailed*
  ai/lzy/v1/channel/channel-manager-private.proto:28-84    GetOrCreate*, Destroy*, DestroyAll*, Status*, List*
  ai/lzy/v1/common/env.proto:10-52             EnvSpec, PythonEnv, ProcessEnv, LocalModule, DockerCredentials, DockerPullPolicy
  ai/lzy/v1/common/slot.proto:10-29            Slot{Media, Direction}
  ai/lzy/v1/common/operation.proto:12-45       Operation, Requirements, SlotToChannelAssignment, TaskDesc, KafkaTopicDescription
  ai/lzy/v1/worker/worker-service.proto:14-50  WorkerApi: Init*, Execute*
  ai/lzy/v1/long-running/operation.proto:12-47 LongRunningService.Get, Operation (google.rpc.Status restated: code, message, details)

and one additive service of this repository's own (package ai.lzy.b200.v1, nothing of the reference's changes):
  LzybPeerApi.Describe(peer_id) -> the 256-byte lzyb_ipc_desc_t of a producer slot (include/lzy_b200_host.h: lzyb_slots_describe)

A message serialized here parses with the reference's generated classes and vice versa (same
package, names, numbers and types)."""
from __future__ import annotations

from typing import Dict, Tuple

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

F = descriptor_pb2.FieldDescriptorProto
_pool = descriptor_pool.DescriptorPool()


def _field(name, number, ftype, type_name=None, label=F.LABEL_OPTIONAL, oneof=None):
    f = F(name=name, number=number, type=ftype, label=label)
    if type_name:
        f.type_name = type_name
    if oneof is not None:
        f.oneof_index = oneof
    return f


def _msg(name, fields, oneofs=(), nested=(), enums=()):
    m = descriptor_pb2.DescriptorProto(name=name)
    m.field.extend(fields)
    for o in oneofs:
        m.oneof_decl.add(name=o)
    m.nested_type.extend(nested)
    m.enum_type.extend(enums)
    return m


def _enum(name, values):
    e = descriptor_pb2.EnumDescriptorProto(name=name)
    for n, v in values:
        e.value.add(name=n, number=v)
    return e


STR, U64, BYTES, MSG, ENUM = F.TYPE_STRING, F.TYPE_UINT64, F.TYPE_BYTES, F.TYPE_MESSAGE, F.TYPE_ENUM
REP = F.LABEL_REPEATED

# ---- ai.lzy.v1.common ---------------------------------------------------------------------------------
_common = descriptor_pb2.FileDescriptorProto(name="lzyb/ai/lzy/v1/common/common.proto", package="ai.lzy.v1.common",
                                              syntax="proto3")
_meta_entry = _msg("MetadataEntry", [_field("key", 1, STR), _field("value", 2, STR)])
_meta_entry.options.map_entry = True
_common.message_type.extend([
    _msg("DataScheme", [_field("dataFormat", 1, STR), _field("schemeFormat", 2, STR), _field("schemeContent", 3, STR),
                        _field("metadata", 10, MSG, ".ai.lzy.v1.common.DataScheme.MetadataEntry", REP)], nested=[_meta_entry]),
    _msg("S3Credentials", [_field("endpoint", 1, STR), _field("accessToken", 2, STR), _field("secretToken", 3, STR)]),
    _msg("AzureBlobStorageCredentials", [_field("connectionString", 1, STR)]),
    _msg("PeerDescription",
         [_field("peer_id", 1, STR),
          _field("slot_peer", 2, MSG, ".ai.lzy.v1.common.PeerDescription.SlotPeer", oneof=0),
          _field("storage_peer", 3, MSG, ".ai.lzy.v1.common.PeerDescription.StoragePeer", oneof=0)],
         oneofs=["peer"],
         nested=[_msg("SlotPeer", [_field("peer_url", 2, STR)]),
                 _msg("StoragePeer", [_field("storage_uri", 1, STR),
                                      _field("s3", 2, MSG, ".ai.lzy.v1.common.S3Credentials", oneof=0),
                                      _field("azure", 3, MSG, ".ai.lzy.v1.common.AzureBlobStorageCredentials", oneof=0)],
                      oneofs=["credentials"])]),
    _msg("ChannelStatus", [_field("channel_id", 1, STR), _field("scheme", 2, MSG, ".ai.lzy.v1.common.DataScheme"),
                           _field("execution_id", 3, STR),
                           _field("consumers", 4, MSG, ".ai.lzy.v1.common.PeerDescription", REP),
                           _field("producers", 5, MSG, ".ai.lzy.v1.common.PeerDescription", REP)]),
])
_pool.Add(_common)

# ---- ai.lzy.v1.slots -----------------------------------------------------------------------------------
_slots = descriptor_pb2.FileDescriptorProto(name="lzyb/ai/lzy/v1/slots/slots-api.proto", package="ai.lzy.v1.slots",
                                             syntax="proto3", dependency=[_common.name])
_slots.message_type.extend([
    _msg("StartTransferRequest", [_field("transfer_id", 1, STR), _field("slot_id", 2, STR),
                                  _field("peer", 3, MSG, ".ai.lzy.v1.common.PeerDescription")]),
    _msg("StartTransferResponse", []),
    _msg("ReadDataRequest", [_field("peer_id", 1, STR), _field("offset", 2, U64)]),
    _msg("ReadDataChunk", [_field("control", 1, ENUM, ".ai.lzy.v1.slots.ReadDataChunk.Control", oneof=0),
                           _field("chunk", 2, BYTES, oneof=0)], oneofs=["kind"], enums=[_enum("Control", [("EOS", 0)])]),
])
_pool.Add(_slots)

# ---- ai.lzy.v1.channel ----------------------------------------------------------------------------------
_chan = descriptor_pb2.FileDescriptorProto(name="lzyb/ai/lzy/v1/channel/channel-manager.proto", package="ai.lzy.v1.channel",
                                            syntax="proto3", dependency=[_common.name])
PD, DS, CS = ".ai.lzy.v1.common.PeerDescription", ".ai.lzy.v1.common.DataScheme", ".ai.lzy.v1.common.ChannelStatus"
SP = ".ai.lzy.v1.common.PeerDescription.StoragePeer"
_chan.message_type.extend([
    _msg("BindRequest", [_field("execution_id", 1, STR), _field("channel_id", 2, STR), _field("peer_id", 3, STR),
                         _field("peer_url", 4, STR), _field("role", 5, ENUM, ".ai.lzy.v1.channel.BindRequest.Role")],
         enums=[_enum("Role", [("UNSPECIFIED", 0), ("CONSUMER", 1), ("PRODUCER", 2)])]),
    _msg("BindResponse", [_field("data_scheme", 1, MSG, DS), _field("peer", 2, MSG, PD), _field("transfer_id", 3, STR)]),
    _msg("UnbindRequest", [_field("peer_id", 1, STR), _field("channel_id", 2, STR)]),
    _msg("UnbindResponse", []),
    _msg("GetChannelsStatusRequest", [_field("execution_id", 1, STR), _field("channel_ids", 2, STR, label=REP)]),
    _msg("GetChannelsStatusResponse", [_field("channels", 1, MSG, CS, REP)]),
    _msg("TransferCompletedRequest", [_field("transfer_id", 1, STR), _field("channel_id", 2, STR)]),
    _msg("TransferCompletedResponse", []),
    _msg("TransferFailedRequest", [_field("transfer_id", 1, STR), _field("channel_id", 2, STR), _field("description", 3, STR)]),
    _msg("TransferFailedResponse", [_field("new_transfer_id", 1, STR), _field("new_peer", 2, MSG, PD)]),
    # private API
    _msg("GetOrCreateRequest", [_field("execution_id", 1, STR), _field("workflow_name", 2, STR), _field("user_id", 3, STR),
                                _field("scheme", 4, MSG, DS), _field("producer", 5, MSG, SP, oneof=0),
                                _field("consumer", 6, MSG, SP, oneof=0)], oneofs=["initial_storage_peer"]),
    _msg("GetOrCreateResponse", [_field("channel_id", 1, STR)]),
    _msg("DestroyRequest", [_field("channel_id", 1, STR), _field("reason", 2, STR)]),
    _msg("DestroyResponse", []),
    _msg("DestroyAllRequest", [_field("execution_id", 1, STR), _field("reason", 2, STR)]),
    _msg("DestroyAllResponse", []),
    _msg("StatusRequest", [_field("channel_id", 1, STR)]),
    _msg("StatusResponse", [_field("channel", 1, MSG, CS)]),
    _msg("ListRequest", [_field("execution_id", 1, STR)]),
    _msg("ListResponse", [_field("statuses", 1, MSG, CS, REP)]),
])
_pool.Add(_chan)


# ---- ai.lzy.v1.common: what a task description carries ----------------------------------------------------
I32, I64, BOOL = F.TYPE_INT32, F.TYPE_INT64, F.TYPE_BOOL
_task = descriptor_pb2.FileDescriptorProto(name="lzyb/ai/lzy/v1/common/operation.proto", package="ai.lzy.v1.common",
                                            syntax="proto3", dependency=[_common.name])
_env_entry = _msg("EnvEntry", [_field("key", 1, STR), _field("value", 2, STR)])
_env_entry.options.map_entry = True
_task.enum_type.extend([_enum("DockerPullPolicy", [("UNSPECIFIED", 0), ("ALWAYS", 1), ("IF_NOT_EXISTS", 2)])])
_task.message_type.extend([
    _msg("DockerCredentials", [_field("registry_name", 1, STR), _field("username", 2, STR), _field("password", 3, STR)]),
    _msg("LocalModule", [_field("name", 1, STR), _field("uri", 2, STR)]),
    _msg("PythonEnv", [_field("name", 1, STR), _field("yaml", 2, STR),
                       _field("local_modules", 3, MSG, ".ai.lzy.v1.common.LocalModule", REP)]),
    _msg("ProcessEnv", []),
    _msg("EnvSpec", [_field("docker_pull_policy", 1, ENUM, ".ai.lzy.v1.common.DockerPullPolicy"),
                     _field("docker_credentials", 2, MSG, ".ai.lzy.v1.common.DockerCredentials"),
                     _field("docker_image", 3, STR),
                     _field("env", 4, MSG, ".ai.lzy.v1.common.EnvSpec.EnvEntry", REP),
                     _field("pyenv", 5, MSG, ".ai.lzy.v1.common.PythonEnv", oneof=0),
                     _field("process_env", 6, MSG, ".ai.lzy.v1.common.ProcessEnv", oneof=0)],
         oneofs=["execution_context"], nested=[_env_entry]),
    _msg("Slot", [_field("name", 1, STR), _field("media", 2, ENUM, ".ai.lzy.v1.common.Slot.Media"),
                  _field("contentType", 3, MSG, DS), _field("direction", 4, ENUM, ".ai.lzy.v1.common.Slot.Direction")],
         enums=[_enum("Media", [("UNSPECIFIED", 0), ("FILE", 1), ("PIPE", 2), ("ARG", 3)]),
                _enum("Direction", [("UNKNOWN", 0), ("INPUT", 1), ("OUTPUT", 2)])]),
    _msg("KafkaTopicDescription", [_field("bootstrap_servers", 1, STR, label=REP), _field("username", 2, STR),
                                   _field("password", 3, STR), _field("topic", 4, STR)]),
    _msg("Requirements", [_field("pool_label", 1, STR), _field("zone", 2, STR)]),
    _msg("Operation", [_field("env", 1, MSG, ".ai.lzy.v1.common.EnvSpec"),
                       _field("requirements", 2, MSG, ".ai.lzy.v1.common.Requirements"), _field("command", 3, STR),
                       _field("slots", 4, MSG, ".ai.lzy.v1.common.Slot", REP), _field("description", 5, STR),
                       _field("name", 6, STR), _field("kafka_topic", 9, MSG, ".ai.lzy.v1.common.KafkaTopicDescription")]),
    _msg("SlotToChannelAssignment", [_field("slot_name", 1, STR), _field("channel_id", 2, STR)]),
    _msg("TaskDesc", [_field("operation", 2, MSG, ".ai.lzy.v1.common.Operation"),
                      _field("slot_assignments", 3, MSG, ".ai.lzy.v1.common.SlotToChannelAssignment", REP)]),
])
_pool.Add(_task)

# ---- google.rpc.Status + ai.lzy.v1.longrunning ---------------------------------------------------------------
from google.protobuf import any_pb2, timestamp_pb2  # noqa: E402

for _wkt in (any_pb2, timestamp_pb2):
    _pool.AddSerializedFile(_wkt.DESCRIPTOR.serialized_pb)
_rpc = descriptor_pb2.FileDescriptorProto(name="lzyb/google/rpc/status.proto", package="google.rpc", syntax="proto3",
                                           dependency=["google/protobuf/any.proto"])
_rpc.message_type.extend([_msg("Status", [_field("code", 1, I32), _field("message", 2, STR),
                                          _field("details", 3, MSG, ".google.protobuf.Any", REP)])])
_pool.Add(_rpc)
_lro = descriptor_pb2.FileDescriptorProto(name="lzyb/ai/lzy/v1/long-running/operation.proto", package="ai.lzy.v1.longrunning",
                                           syntax="proto3",
                                           dependency=["google/protobuf/any.proto", "google/protobuf/timestamp.proto", _rpc.name])
_lro.message_type.extend([
    _msg("Operation", [_field("id", 1, STR), _field("description", 2, STR),
                       _field("created_at", 3, MSG, ".google.protobuf.Timestamp"), _field("created_by", 4, STR),
                       _field("modified_at", 5, MSG, ".google.protobuf.Timestamp"), _field("done", 6, BOOL),
                       _field("metadata", 7, MSG, ".google.protobuf.Any"),
                       _field("error", 8, MSG, ".google.rpc.Status", oneof=0),
                       _field("response", 9, MSG, ".google.protobuf.Any", oneof=0)], oneofs=["result"]),
    _msg("GetOperationRequest", [_field("operation_id", 1, STR)]),
])
_pool.Add(_lro)

# ---- ai.lzy.v1.worker ----------------------------------------------------------------------------------------
_worker = descriptor_pb2.FileDescriptorProto(name="lzyb/ai/lzy/v1/worker/worker-service.proto", package="ai.lzy.v1.worker",
                                              syntax="proto3", dependency=[_task.name, _lro.name])
_worker.message_type.extend([
    _msg("InitRequest", [_field("user_id", 1, STR), _field("workflow_name", 2, STR), _field("worker_subject_name", 3, STR),
                         _field("worker_private_key", 4, STR)]),
    _msg("InitResponse", []),
    _msg("ExecuteRequest", [_field("task_desc", 1, MSG, ".ai.lzy.v1.common.TaskDesc"), _field("task_id", 2, STR),
                            _field("execution_id", 3, STR), _field("workflow_name", 4, STR), _field("user_id", 5, STR)]),
    _msg("ExecuteResponse", [_field("rc", 1, I32), _field("description", 2, STR)]),
])
_pool.Add(_worker)


# ---- ai.lzy.b200.v1: the one additive RPC (peers in other worker processes) ------------------------------------
_peer = descriptor_pb2.FileDescriptorProto(name="lzyb/ai/lzy/b200/v1/peer-api.proto", package="ai.lzy.b200.v1", syntax="proto3")
_peer.message_type.extend([
    _msg("DescribeRequest", [_field("peer_id", 1, STR)]),
    _msg("DescribeResponse", [_field("descriptor", 1, BYTES)]),
])
_pool.Add(_peer)


def _cls(full_name):
    return message_factory.GetMessageClass(_pool.FindMessageTypeByName(full_name))


class _Messages:
    """attribute access to message classes: ``messages.BindRequest``, ``messages.PeerDescription`` …"""
    _packages = ("ai.lzy.v1.common", "ai.lzy.v1.slots", "ai.lzy.v1.channel", "ai.lzy.v1.worker")

    def __getattr__(self, name):
        for pkg in self._packages:
            try:
                cls = _cls(f"{pkg}.{name}")
            except KeyError:
                continue
            setattr(self, name, cls)
            return cls
        raise AttributeError(name)


messages = _Messages()
LongRunningOperation = _cls("ai.lzy.v1.longrunning.Operation")   # (common also has an `Operation`: the task's)
GetOperationRequest = _cls("ai.lzy.v1.longrunning.GetOperationRequest")
RpcStatus = _cls("google.rpc.Status")
DescribeRequest = _cls("ai.lzy.b200.v1.DescribeRequest")
DescribeResponse = _cls("ai.lzy.b200.v1.DescribeResponse")

# service → method → (request class name, response class name, server-streaming?)
SERVICES: Dict[str, Dict[str, Tuple[str, str, bool]]] = {
    "ai.lzy.v1.slots.LzySlotsApi": {
        "StartTransfer": ("StartTransferRequest", "StartTransferResponse", False),
        "Read": ("ReadDataRequest", "ReadDataChunk", True),
    },
    "ai.lzy.v1.channel.LzyChannelManager": {
        "Bind": ("BindRequest", "BindResponse", False),
        "Unbind": ("UnbindRequest", "UnbindResponse", False),
        "TransferCompleted": ("TransferCompletedRequest", "TransferCompletedResponse", False),
        "TransferFailed": ("TransferFailedRequest", "TransferFailedResponse", False),
        "GetChannelsStatus": ("GetChannelsStatusRequest", "GetChannelsStatusResponse", False),
    },
    "ai.lzy.v1.channel.LzyChannelManagerPrivate": {
        "GetOrCreate": ("GetOrCreateRequest", "GetOrCreateResponse", False),
        "Destroy": ("DestroyRequest", "DestroyResponse", False),
        "DestroyAll": ("DestroyAllRequest", "DestroyAllResponse", False),
        "Status": ("StatusRequest", "StatusResponse", False),
        "List": ("ListRequest", "ListResponse", False),
    },
    "ai.lzy.v1.worker.WorkerApi": {
        "Init": ("InitRequest", "InitResponse", False),
        "Execute": ("ExecuteRequest", "ai.lzy.v1.longrunning.Operation", False),
    },
    "ai.lzy.b200.v1.LzybPeerApi": {
        "Describe": ("ai.lzy.b200.v1.DescribeRequest", "ai.lzy.b200.v1.DescribeResponse", False),
    },
    "ai.lzy.v1.longrunning.LongRunningService": {
        "Get": ("ai.lzy.v1.longrunning.GetOperationRequest", "ai.lzy.v1.longrunning.Operation", False),
    },
}


def method_types(service: str, method: str):
    req, resp, streaming = SERVICES[service][method]
    resolve = lambda n: _cls(n) if "." in n else getattr(messages, n)  # noqa: E731
    return resolve(req), resolve(resp), streaming
