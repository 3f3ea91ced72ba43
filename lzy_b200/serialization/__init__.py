from .api import Schema, Serializer, SerializerRegistry, StandardDataFormats, StandardSchemaFormats
from .registry import LzySerializerRegistry, SerializerImport
from .types import EmptyContent, get_type

__all__ = ["Schema", "Serializer", "SerializerRegistry", "StandardDataFormats", "StandardSchemaFormats",
           "LzySerializerRegistry", "SerializerImport", "EmptyContent", "get_type"]
