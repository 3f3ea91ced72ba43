"""Serializer interface and registry.

Shape of the third-party ``serialzy`` API the reference builds on (visible through
pylzy/lzy/serialization/file.py:16-87, pylzy/lzy/serialization/registry.py:20-82 and
pylzy/tests/scenarios/custom_serializer/serializer.py:9-33): a serializer says which types it
handles, writes an object to a binary stream with **no framing of its own added by the base class**,
reads it back, and describes the data with a ``Schema``.  When the real ``serialzy`` is importable
these classes are still used — user serializers written against serialzy's ``Serializer`` are accepted
by duck typing (same method names).

On top of that byte-stream contract a serializer may implement the *device-frame* extension
(``DeviceFrameSerializer``): the payload is packed straight into an HBM slot frame by the CUDA pack
kernel and never becomes a Python ``bytes`` object.
"""
from __future__ import annotations

import abc
import enum
from dataclasses import dataclass, field
from typing import Any, BinaryIO, Callable, Dict, List, Optional, Type, Union


class StandardDataFormats(enum.Enum):
    pickle = "pickle"
    proto = "proto"
    raw_type = "raw_type"
    primitive_type = "primitive_type"


class StandardSchemaFormats(enum.Enum):
    pickled_type = "pickled_type"
    json_pickled_type = "json_pickled_type"
    no_schema = "no_schema"


@dataclass(frozen=True)
class VersionBoundary:
    greater_than: Optional[str] = None
    less_than: Optional[str] = None


@dataclass
class Schema:
    """LMD.DataScheme (model/src/main/proto/ai/lzy/v1/common/data-scheme.proto:8-14)"""
    data_format: str
    schema_format: str
    schema_content: str = ""
    meta: Dict[str, str] = field(default_factory=dict)


class Serializer(abc.ABC):
    @abc.abstractmethod
    def supported_types(self) -> Union[Type, Callable[[Type], bool]]:
        ...

    @abc.abstractmethod
    def data_format(self) -> str:
        ...

    @abc.abstractmethod
    def _serialize(self, obj: Any, dest: BinaryIO) -> None:
        ...

    @abc.abstractmethod
    def _deserialize(self, source: BinaryIO, schema_type: Type, user_type: Optional[Type] = None) -> Any:
        ...

    def available(self) -> bool:
        return True

    def stable(self) -> bool:
        return False

    def meta(self) -> Dict[str, str]:
        return {}

    def requirements(self) -> Dict[str, VersionBoundary]:
        return {}

    def schema(self, typ: type) -> Schema:
        return Schema(self.data_format(), StandardSchemaFormats.no_schema.name, meta=self.meta())

    def resolve(self, schema: Schema) -> Type:
        if schema.data_format != self.data_format():
            raise ValueError(f"Invalid data format {schema.data_format}")
        t = self.supported_types()
        return t if isinstance(t, type) else object

    # public entry points (no framing is added around _serialize's bytes)
    def serialize(self, obj: Any, dest: BinaryIO) -> None:
        self._serialize(obj, dest)

    def deserialize(self, source: BinaryIO, typ: Optional[Type] = None) -> Any:
        t = self.supported_types()
        return self._deserialize(source, t if isinstance(t, type) else object, typ)

    def _check_types_valid(self, schema_type: Type, user_type: Optional[Type]) -> None:
        if user_type is not None and isinstance(user_type, type) and isinstance(schema_type, type):
            if not (issubclass(schema_type, user_type) or issubclass(user_type, schema_type)):
                raise TypeError(f"Cannot deserialize data with schema type {schema_type} into type {user_type}")

    def handles(self, typ: Type) -> bool:
        t = self.supported_types()
        if isinstance(t, type):
            return typ == t
        return bool(t(typ))


class DeviceFrameSerializer(Serializer):
    """Extension: the payload lives in an HBM slot frame.

    ``frame_plan(obj)`` describes the frame (dtype code, shape, byte count); ``pack(obj, slot, stream)``
    fills a fresh output slot (``lzyb_slot_pack*``); ``unpack(slot, header, gpu)`` rebuilds the object
    from a delivered frame."""

    @abc.abstractmethod
    def frame_plan(self, obj: Any):
        ...

    @abc.abstractmethod
    def pack(self, obj: Any, slot, stream: int) -> None:
        ...

    @abc.abstractmethod
    def unpack(self, slot, info, private: bool) -> Any:
        ...


class SerializerRegistry:
    """Priority-ordered registry (serialzy.registry.DefaultSerializerRegistry's role)."""

    def __init__(self):
        self._serializers: List[Serializer] = []
        self._serializer_priorities: Dict[Type[Serializer], int] = {}
        self._next_default = 1000
        # lookups sit on the per-op path (every argument and result asks twice): remembered per type / format until the
        # registry changes
        self._by_type: Dict[Any, Optional[Serializer]] = {}
        self._by_format: Dict[str, Optional[Serializer]] = {}

    def register_serializer(self, serializer: Serializer, priority: Optional[int] = None) -> None:
        if priority is None:
            priority = self._next_default
            self._next_default += 1
        for s in self._serializers:
            if type(s) is type(serializer):
                raise ValueError(f"Serializer {type(serializer).__name__} has been already registered")
        self._serializer_priorities[type(serializer)] = priority
        self._serializers.append(serializer)
        self._serializers.sort(key=lambda s: self._serializer_priorities[type(s)])
        self._by_type.clear()
        self._by_format.clear()

    def unregister_serializer(self, serializer: Serializer) -> None:
        self._serializers = [s for s in self._serializers if type(s) is not type(serializer)]
        self._serializer_priorities.pop(type(serializer), None)
        self._by_type.clear()
        self._by_format.clear()

    def find_serializer_by_type(self, typ: Type) -> Optional[Serializer]:
        try:
            return self._by_type[typ]
        except KeyError:
            pass
        except TypeError:  # an unhashable annotation object: look it up the long way every time
            return self._scan_by_type(typ)
        found = self._scan_by_type(typ)
        self._by_type[typ] = found
        return found

    def _scan_by_type(self, typ: Type) -> Optional[Serializer]:
        for s in self._serializers:
            try:
                if s.handles(typ):
                    return s
            except Exception:
                continue
        return None

    def find_serializer_by_instance(self, obj: Any) -> Optional[Serializer]:
        return self.find_serializer_by_type(type(obj))

    def find_serializer_by_data_format(self, data_format: str) -> Optional[Serializer]:
        try:
            return self._by_format[data_format]
        except KeyError:
            pass
        found = None
        for s in self._serializers:
            if s.data_format() == data_format:
                found = s
                break
        self._by_format[data_format] = found
        return found

    def resolve(self, schema: Schema) -> Type:
        s = self.find_serializer_by_data_format(schema.data_format)
        if s is None:
            raise ValueError(f"Cannot find serializer for data format {schema.data_format}")
        return s.resolve(schema)
