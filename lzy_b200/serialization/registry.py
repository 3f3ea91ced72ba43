"""``lzy.serializer_registry`` — what pylzy/lzy/serialization/registry.py:20-82 gives its users, over this package's
serializers:

* the built-ins (device frames for tensors / ndarrays, ``File``, cloudpickle as the last resort);
* user serializers, remembered by class so that the op side can be told to import them
  (``imports()`` → ``SerializerImport(module, class, priority)`` → ``load_imports()``), and refused when they live in
  ``__main__`` (the op side could not import them);
* values of a registered type *nested* inside a pickled object go through their serializer too, by way of
  ``copyreg`` — for user serializers and ``File``; arrays inside a pickle instead ride as protocol-5 out-of-band
  buffers of one composite frame, so the device-frame serializers are kept out of ``copyreg``."""
from __future__ import annotations

import copyreg
import importlib
from dataclasses import dataclass
from io import BytesIO
from typing import Any, Dict, List, Optional, Sequence, Type

from .api import Serializer, SerializerRegistry
from .builtin import CloudpickleSerializer, FileSerializer, NdarrayFrameSerializer, TensorFrameSerializer

_BUILTINS = ((TensorFrameSerializer, 10), (NdarrayFrameSerializer, 11), (FileSerializer, 20),
             (CloudpickleSerializer, 1_000_000))  # the fallback always comes last
_FIRST_USER_PRIORITY = 100  # user serializers without a priority line up here: before the pickle fallback


@dataclass
class SerializerImport:
    module_name: str
    class_name: str
    priority: int


def _exact_type_of(serializer: Serializer) -> Optional[type]:
    """the one class a by-type serializer stands for (None for predicate serializers or missing libraries)"""
    try:
        supported = serializer.supported_types()
    except (ImportError, ModuleNotFoundError):
        return None
    return supported if isinstance(supported, type) else None


class LzySerializerRegistry(SerializerRegistry):
    def __init__(self):
        super().__init__()
        self._user_classes: Dict[Type[Serializer], None] = {}  # insertion-ordered set
        self._open_for_users = False
        for cls, priority in _BUILTINS:
            self.register_serializer(cls(), priority)
        self._open_for_users = True

    # ---- registration -----------------------------------------------------------------------------------------------
    def register_serializer(self, serializer: Serializer, priority: Optional[int] = None) -> None:
        cls = type(serializer)
        if cls.__module__ == "__main__":
            raise ValueError("Cannot register serializers from the __main__ module. "
                             "Please move serializer class to an external file.")
        from_user = self._open_for_users
        if from_user and priority is None:
            priority = _FIRST_USER_PRIORITY + len(self._user_classes)
        super().register_serializer(serializer, priority)
        if from_user:
            self._user_classes[cls] = None
        if from_user or isinstance(serializer, FileSerializer):
            typ = _exact_type_of(serializer)
            if typ is not None:
                copyreg.dispatch_table[typ] = self._reduce_nested

    def unregister_serializer(self, serializer: Serializer) -> None:
        super().unregister_serializer(serializer)
        self._user_classes.pop(type(serializer), None)
        typ = _exact_type_of(serializer)
        if typ is not None:
            copyreg.dispatch_table.pop(typ, None)

    # ---- shipping user serializers to the op side -----------------------------------------------------------------
    def imports(self) -> Sequence[SerializerImport]:
        shipped: List[SerializerImport] = []
        for cls in self._user_classes:
            shipped.append(SerializerImport(cls.__module__, cls.__name__, self._serializer_priorities[cls]))
        return shipped

    def load_imports(self, imports: Sequence[SerializerImport]) -> None:
        for wanted in imports:
            cls = getattr(importlib.import_module(wanted.module_name), wanted.class_name)
            if cls not in self._serializer_priorities:  # already here (same process, or loaded twice)
                self.register_serializer(cls(), wanted.priority)

    # ---- nested values inside pickles -------------------------------------------------------------------------------
    def _reduce_nested(self, obj: Any):
        serializer = self.find_serializer_by_instance(obj)
        raw = BytesIO()
        serializer.serialize(obj, raw)  # type: ignore[union-attr]
        return self._revive_nested, (serializer.data_format(), raw.getvalue())  # type: ignore[union-attr]

    def _revive_nested(self, data_format: str, raw: bytes) -> Any:
        serializer = self.find_serializer_by_data_format(data_format)
        return serializer.deserialize(BytesIO(raw))  # type: ignore[union-attr]
