"""Protobuf messages of the kept wire contract, declared programmatically (this image has no
``protoc``/``grpc_tools``).  Field names and numbers restate, message by message:

  ai/lzy/v1/common/data-scheme.proto:8-14      DataScheme
  ai/lzy/v1/common/storage.proto:10-18         S3Credentials, AzureBlobStorageCredentials
  ai/lzy/v1/common/channel.proto:12-40         ChannelStatus, PeerDescription{SlotPeer, StoragePeer}
  ai/lzy/v1/slots/slots-api.proto:23-47        StartTransferRequest/Response, ReadDataRequest, ReadDataChunk
  ai/lzy/v1/channel/channel-manager.proto:27-95            Bind*, Unbind*, GetChannelsStatus*, TransferCompleted*, TransferFailed*
  ai/lzy/v1/channel/channel-manager-private.proto:28-84    GetOrCreate*, Destroy*, DestroyAll*, Status*, List*

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


def _cls(full_name):
    return message_factory.GetMessageClass(_pool.FindMessageTypeByName(full_name))


class _Messages:
    """attribute access to message classes: ``messages.BindRequest``, ``messages.PeerDescription`` …"""
    _packages = ("ai.lzy.v1.common", "ai.lzy.v1.slots", "ai.lzy.v1.channel")

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
}


def method_types(service: str, method: str):
    req, resp, streaming = SERVICES[service][method]
    return getattr(messages, req), getattr(messages, resp), streaming
