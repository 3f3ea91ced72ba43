"""Type helpers for @op signatures (pylzy/lzy/api/v1/utils/types.py:23-79)."""
from __future__ import annotations

import inspect
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Type, Union, get_type_hints

from beartype.door import is_subhint
from typing_extensions import get_args, get_origin

from ....proxy_result import Absence, Either, Result
from ....serialization.api import SerializerRegistry
from ....serialization.types import EmptyContent


def infer_real_types(typ: Type) -> Sequence[Type]:
    origin: Optional[Type] = get_origin(typ)
    result: List[Type] = []
    if origin is not None:
        if origin == Union:  # type: ignore[comparison-overlap]
            for arg in get_args(typ):
                result.extend(infer_real_types(arg))
            return result
        return (origin,)
    return (typ,)


def infer_return_type(func: Callable) -> Either:
    hints = get_type_hints(func)
    if "return" not in hints:
        return Absence()
    typ = hints["return"]
    if isinstance(typ, tuple):  # `-> (A, B)`: several outputs
        return Result(typ)
    return Result((typ,))


def get_default_args(func: Callable) -> Dict[str, Any]:
    signature = inspect.signature(func)
    return {k: v.default for k, v in signature.parameters.items() if v.default is not inspect.Parameter.empty}


def check_types_serialization_compatible(annotation: Type, typ: Type, registry: SerializerRegistry) -> bool:
    if annotation == typ:
        return True
    for annotation_type in infer_real_types(annotation):
        serializer = registry.find_serializer_by_type(annotation_type)
        if serializer is None:
            continue
        supported = serializer.supported_types()
        ok = supported == typ if isinstance(supported, type) else supported(typ)
        if ok:
            return True
    return False


def is_subtype(subtype: Type, supertype: Type) -> bool:
    if subtype == List[EmptyContent]:
        subtype = supertype if is_subhint(supertype, List) else List  # type: ignore[assignment]
    elif subtype == Tuple[EmptyContent]:
        subtype = supertype if is_subhint(supertype, Tuple) else Tuple  # type: ignore[assignment]
    elif subtype == Dict[EmptyContent, EmptyContent]:
        subtype = supertype if is_subhint(supertype, Dict) else Dict  # type: ignore[assignment]
    return is_subhint(subtype, supertype)
