"""LzySerializerRegistry — the registry users see on ``lzy.serializer_registry``
(pylzy/lzy/serialization/registry.py:20-82): built-ins, user serializers tracked so that they can be
shipped to the op side as (module, class, priority) imports, the ``__main__`` ban, and copyreg hooks
so that values of registered types nested inside pickled objects go through their serializer."""
from __future__ import annotations

import copyreg
import importlib
from dataclasses import dataclass
from io import BytesIO
from typing import Any, Optional, Sequence, Set, Type

from .api import Serializer, SerializerRegistry
from .builtin import CloudpickleSerializer, FileSerializer, NdarrayFrameSerializer, TensorFrameSerializer


@dataclass
class SerializerImport:
    module_name: str
    class_name: str
    priority: int


class LzySerializerRegistry(SerializerRegistry):
    def __init__(self):
        super().__init__()
        self.__inited = False
        self.__user_serializers: Set[Type[Serializer]] = set()
        self.register_serializer(TensorFrameSerializer(), 10)
        self.register_serializer(NdarrayFrameSerializer(), 11)
        self.register_serializer(FileSerializer(), 20)
        self.register_serializer(CloudpickleSerializer(), 1_000_000)  # the fallback always comes last
        self.__inited = True

    def register_serializer(self, serializer: Serializer, priority: Optional[int] = None) -> None:
        if type(serializer).__module__ == "__main__":
            raise ValueError("Cannot register serializers from the __main__ module. "
                             "Please move serializer class to an external file.")
        if priority is None and self.__dict__.get("_LzySerializerRegistry__inited"):
            priority = 100 + len(self.__user_serializers)  # user serializers win over the pickle fallback
        super().register_serializer(serializer, priority)
        user = bool(self.__dict__.get("_LzySerializerRegistry__inited"))
        if user:
            self.__user_serializers.add(type(serializer))
        # an instance of a serializer's type nested inside a pickled object travels through that serializer
        # (registry.py:34-46 hooks every by-type serializer into copyreg).  The device-frame serializers stay
        # out: arrays inside a pickle ride as protocol-5 out-of-band buffers of one composite frame instead.
        if user or isinstance(serializer, FileSerializer):
            try:
                t = serializer.supported_types()
                if isinstance(t, type):
                    copyreg.dispatch_table[t] = self.__reducer
            except (ImportError, ModuleNotFoundError):
                pass

    def unregister_serializer(self, serializer: Serializer) -> None:
        super().unregister_serializer(serializer)
        self.__user_serializers.discard(type(serializer))
        try:
            t = serializer.supported_types()
            if isinstance(t, type):
                copyreg.dispatch_table.pop(t, None)
        except (ImportError, ModuleNotFoundError):
            pass

    def imports(self) -> Sequence[SerializerImport]:
        return [SerializerImport(s.__module__, s.__name__, self._serializer_priorities[s]) for s in self.__user_serializers]

    def load_imports(self, imports: Sequence[SerializerImport]) -> None:
        for imp in imports:
            module = importlib.import_module(imp.module_name)
            cls = getattr(module, imp.class_name)
            if cls not in self._serializer_priorities:
                self.register_serializer(cls(), imp.priority)

    def __constructor(self, data_format: str, data: bytes) -> Any:
        serializer = self.find_serializer_by_data_format(data_format)
        return serializer.deserialize(BytesIO(data))  # type: ignore[union-attr]

    def __reducer(self, obj: Any):
        serializer = self.find_serializer_by_instance(obj)
        buffer = BytesIO()
        serializer.serialize(obj, buffer)  # type: ignore[union-attr]
        return self.__constructor, (serializer.data_format(), buffer.getvalue())  # type: ignore[union-attr]
