"""`FileSerializer` — the one serializer whose bytes the reference itself defines (SURVEY §8 row a8,
pylzy/lzy/serialization/file.py:16-87): ``b"permissions"`` + ``oct(mode & 0o777)`` + the file's bytes, with a
header-less stream still accepted.  Scenarios of pylzy/tests/serialization/test_file.py:32-126, plus the exact
bytes."""
import logging
import os
import stat
import tempfile
from dataclasses import dataclass
from io import BytesIO

import pytest

from lzy_b200.serialization.api import Schema, StandardDataFormats, StandardSchemaFormats
from lzy_b200.serialization.registry import LzySerializerRegistry
from lzy_b200.types import File

CONTENT = "test string"


@pytest.fixture
def registry():
    return LzySerializerRegistry()


@pytest.fixture
def source():
    with tempfile.NamedTemporaryFile() as f:
        f.write(CONTENT.encode())
        f.flush()
        yield f.name


def test_exact_bytes_and_round_trip_keep_the_permission_bits(registry, source):
    os.chmod(source, 0o751)
    ser = registry.find_serializer_by_type(File)
    buf = BytesIO()
    ser.serialize(File(source), buf)
    assert buf.getvalue() == b"permissions" + b"0o751" + CONTENT.encode()
    buf.seek(0)
    back = ser.deserialize(buf, File)
    assert str(back) != source and stat.S_IMODE(back.stat().st_mode) == 0o751
    with back.open() as f:
        assert f.read() == CONTENT
    assert ser.stable() and "pylzy" in ser.meta()
    os.unlink(str(back))


def test_a_stream_written_without_the_header_is_still_a_file(registry):
    ser = registry.find_serializer_by_type(File)
    for payload in (CONTENT.encode(), b"short", b"", b"permission denied is not the header"):
        back = ser.deserialize(BytesIO(payload), File)
        with back.open("rb") as f:
            assert f.read() == payload
        os.unlink(str(back))


def test_schema_and_resolve(registry, caplog):
    ser = registry.find_serializer_by_data_format("raw_file")
    schema = ser.schema(File)
    assert (schema.data_format, schema.schema_format, schema.schema_content) == ("raw_file", StandardSchemaFormats.no_schema.name, "")
    assert "pylzy" in schema.meta
    with pytest.raises(ValueError, match="Invalid type"):
        ser.schema(str)

    assert ser.resolve(Schema("raw_file", StandardSchemaFormats.no_schema.name, meta=ser.meta())) is File
    with pytest.raises(ValueError, match="Invalid data format"):
        ser.resolve(Schema(StandardDataFormats.proto.name, StandardSchemaFormats.no_schema.name, meta={"pylzy": "0.0.0"}))
    with pytest.raises(ValueError, match="Invalid schema format"):
        ser.resolve(Schema("raw_file", StandardSchemaFormats.pickled_type.name, meta={"pylzy": "0.0.0"}))
    with caplog.at_level(logging.WARNING):
        caplog.clear()
        ser.resolve(Schema("raw_file", StandardSchemaFormats.no_schema.name))
        assert caplog.messages and caplog.messages[0].startswith("No pylzy version in meta")
        caplog.clear()
        ser.resolve(Schema("raw_file", StandardSchemaFormats.no_schema.name, meta={"pylzy": "100000.0.0"}))
        assert caplog.messages and caplog.messages[0].startswith("Installed version of pylzy")
        caplog.clear()
        ser.resolve(Schema("raw_file", StandardSchemaFormats.no_schema.name, meta={"pylzy": "0.0.0"}))
        assert not caplog.messages


@dataclass
class Attachment:
    file: File
    note: str = ""


def test_a_file_nested_in_another_object_is_carried_by_value(registry, source):
    """the registry's reducer hook (pylzy/lzy/serialization/registry.py:48-58,73-82): a File field inside a pickled
    object travels as its bytes and comes back as a new local file"""
    ser = registry.find_serializer_by_type(Attachment)
    buf = BytesIO()
    obj = Attachment(File(source), "n")
    ser.serialize(obj, buf)
    buf.seek(0)
    back = ser.deserialize(buf)
    assert back.note == "n" and str(back.file) != str(obj.file)
    with back.file.open() as f:
        assert f.read() == CONTENT
    os.unlink(str(back.file))


def test_raw_file_format_is_the_one_the_reference_serializer_writes_and_reads(tmp_path):
    """tests/golden/reference_file_vectors.json: pylzy/lzy/serialization/file.py's own FileSerializer (executed in the
    authoring container) on fifteen files (content × mode) and seven streams — byte for byte on the way out, content
    and permissions on the way in, headerless and too-short streams included"""
    import importlib.util
    import json
    from pathlib import Path

    from lzy_b200.serialization.builtin import FileSerializer
    from lzy_b200.types import File

    golden = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("file_battery", golden / "file_battery.py")
    battery = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(battery)
    ser = FileSerializer()
    ours = battery.run(dict(File=File, serialize=ser._serialize, deserialize=lambda source: ser._deserialize(source, File)),
                       str(tmp_path))
    assert json.loads(json.dumps(ours)) == json.loads((golden / "reference_file_vectors.json").read_text())
