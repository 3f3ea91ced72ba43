"""lzy_b200/grpc_api/protos.py declares the kept wire contract by hand (the image has no protoc).  This test holds it,
mechanically, to the reference's own .proto files as parsed by tests/golden/make_proto_golden.py: packages, message
names, every field's name / number / type / label / oneof membership, nested types, enum values, services, rpc
signatures and streaming flags.  The three files that ARE the slot/channel surface (slots-api.proto:13-47,
channel-manager.proto:13-95, channel-manager-private.proto:13-84) must be covered completely; of the common files
(everything they import) whatever protos.py declares must match, field for field."""
import json
from pathlib import Path

import pytest
from google.protobuf.descriptor import FieldDescriptor as FD

from lzy_b200.grpc_api import protos

GOLDEN = json.loads((Path(__file__).resolve().parent / "golden" / "reference_proto_vectors.json").read_text())
SURFACE = [f for f in GOLDEN if f.endswith(("slots-api.proto", "channel-manager.proto", "channel-manager-private.proto"))]
SCALARS = {"string": FD.TYPE_STRING, "bytes": FD.TYPE_BYTES, "bool": FD.TYPE_BOOL, "uint64": FD.TYPE_UINT64, "int64": FD.TYPE_INT64,
           "uint32": FD.TYPE_UINT32, "int32": FD.TYPE_INT32, "double": FD.TYPE_DOUBLE, "float": FD.TYPE_FLOAT,
           "sint32": FD.TYPE_SINT32, "sint64": FD.TYPE_SINT64, "fixed64": FD.TYPE_FIXED64, "fixed32": FD.TYPE_FIXED32}


def find(full_name):
    try:
        return protos._pool.FindMessageTypeByName(full_name)
    except KeyError:
        return None


def check_type(field, declared: str, where: str):
    if declared in SCALARS:
        assert field.type == SCALARS[declared], f"{where}: type {field.type}, the .proto says {declared}"
        return
    if declared.startswith("map<"):
        k, v = declared[4:-1].split(",")
        assert field.message_type is not None and field.message_type.GetOptions().map_entry, f"{where}: not a map"
        check_type(field.message_type.fields_by_name["key"], k, where + ".key")
        check_type(field.message_type.fields_by_name["value"], v, where + ".value")
        return
    target = field.message_type or field.enum_type
    assert target is not None, f"{where}: the .proto says {declared}, here it is scalar type {field.type}"
    # the .proto may name the type relative to its scope; the resolved full name must end with what it wrote
    assert ("." + target.full_name).endswith("." + declared), f"{where}: {target.full_name} vs {declared}"


def check_message(pkg: str, name: str, fields, enums):
    full = f"{pkg}.{name}"
    desc = find(full)
    assert desc is not None, f"{full} is not declared"
    assert {f["name"] for f in fields} == {f.name for f in desc.fields}, full
    for f in fields:
        d = desc.fields_by_name[f["name"]]
        where = f"{full}.{f['name']}"
        assert d.number == f["number"], where
        is_repeated = d.is_repeated if hasattr(d, "is_repeated") else d.label == FD.LABEL_REPEATED
        assert is_repeated == (f["label"] == "repeated"), where
        assert (d.containing_oneof.name if d.containing_oneof and not d.containing_oneof.name.startswith("_") else None) == f["oneof"], where
        check_type(d, f["type"], where)
    for ename, values in enums.items():
        if ename.startswith(name + "."):
            e = desc.enum_types_by_name.get(ename[len(name) + 1:])
            if "." in ename[len(name) + 1:]:
                continue  # an enum of a nested message: checked with that message
            assert e is not None, f"{full}: enum {ename} missing"
            assert {v.name: v.number for v in e.values} == values, f"{pkg}.{ename}"


@pytest.mark.parametrize("path", SURFACE)
def test_the_slot_and_channel_surface_is_declared_completely(path):
    spec = GOLDEN[path]
    for name, fields in spec["messages"].items():
        check_message(spec["package"], name, fields, spec["enums"])
    for service, rpcs in spec["services"].items():
        ours = protos.SERVICES.get(f"{spec['package']}.{service}")
        assert ours is not None, service
        assert set(ours) == set(rpcs), service
        for rpc, sig in rpcs.items():
            req, resp, streaming = ours[rpc]
            assert not sig["client_streaming"]
            assert streaming == sig["server_streaming"], f"{service}.{rpc}"
            req_cls, resp_cls, _ = protos.method_types(f"{spec['package']}.{service}", rpc)
            assert ("." + req_cls.DESCRIPTOR.full_name).endswith("." + sig["request"]), f"{service}.{rpc}"
            assert ("." + resp_cls.DESCRIPTOR.full_name).endswith("." + sig["response"]), f"{service}.{rpc}"


def test_everything_declared_from_the_imported_files_matches_them():
    checked = 0
    for path, spec in GOLDEN.items():
        if path in SURFACE:
            continue
        for name, fields in spec["messages"].items():
            if find(f"{spec['package']}.{name}") is None:
                continue  # not part of the kept contract (e.g. Kafka / docker details nobody on this path sends)
            check_message(spec["package"], name, fields, spec["enums"])
            checked += 1
        for service, rpcs in spec["services"].items():
            ours = protos.SERVICES.get(f"{spec['package']}.{service}")
            if ours is None:
                continue
            for rpc, (req, resp, streaming) in ours.items():
                assert rpc in rpcs, f"{service}.{rpc} does not exist in the reference"
                assert streaming == rpcs[rpc]["server_streaming"]
    assert checked >= 12


def test_top_level_enums_match():
    for path, spec in GOLDEN.items():
        for ename, values in spec["enums"].items():
            if "." in ename:
                continue
            try:
                e = protos._pool.FindEnumTypeByName(f"{spec['package']}.{ename}")
            except KeyError:
                continue
            assert {v.name: v.number for v in e.values} == values, ename
