"""A user serializer living outside ``__main__`` (the shape of
pylzy/tests/scenarios/custom_serializer/serializer.py:9-33)."""
from dataclasses import dataclass
from typing import Any, BinaryIO, Callable, Optional, Type, Union

from lzy_b200.serialization.api import Serializer


@dataclass
class Point:
    x: int
    y: int


class PointSerializer(Serializer):
    def supported_types(self) -> Union[Type, Callable[[Type], bool]]:
        return Point

    def data_format(self) -> str:
        return "point_text"

    def stable(self) -> bool:
        return True

    def _serialize(self, obj: Point, dest: BinaryIO) -> None:
        dest.write(f"point:{obj.x},{obj.y}".encode())

    def _deserialize(self, source: BinaryIO, schema_type: Type, user_type: Optional[Type] = None) -> Any:
        x, y = source.read().decode().split(":")[1].split(",")
        return Point(int(x), int(y))
