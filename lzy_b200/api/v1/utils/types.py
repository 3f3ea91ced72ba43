"""What the ``@op`` machinery needs to know about annotations — the same questions pylzy asks in
pylzy/lzy/api/v1/utils/types.py:23-79, answered here over this package's serializer registry:

* which concrete classes may a value declared as ``T`` have (a proxy must answer ``isinstance`` for them),
* what does the function return (one entry, or several for a literal ``-> (A, B)``),
* which parameters have defaults,
* can the serializer chosen for an annotation write a value of some runtime type,
* is a runtime type acceptable for an annotation (with the empty-container rule)."""
from __future__ import annotations

import inspect
from typing import Any, Callable, Dict, List, Sequence, Tuple, Type, Union, get_type_hints

from beartype.door import is_subhint
from typing_extensions import get_args, get_origin

from ....proxy_result import Absence, Either, Result
from ....serialization.api import SerializerRegistry
from ....serialization.types import EmptyContent

_NOT_DECLARED = object()
# get_type() of an empty list / tuple / dict: it fits every annotation of the same container kind
_EMPTY_CONTAINER_KIND = {List[EmptyContent]: List, Tuple[EmptyContent]: Tuple, Dict[EmptyContent, EmptyContent]: Dict}


def infer_real_types(typ: Type) -> Sequence[Type]:
    """``Union[A, List[B]]`` → ``[A, list]``; a generic alias → its origin; a plain class → itself"""
    origin = get_origin(typ)
    if origin is None:
        return (typ,)
    if origin is not Union:
        return (origin,)
    members: List[Type] = []
    for member in get_args(typ):
        members += infer_real_types(member)
    return members


def infer_return_type(func: Callable) -> Either:
    """``Result`` of a tuple of output types — one element unless the annotation is itself a tuple of types
    (``-> (A, B)`` declares two outputs, ``-> Tuple[A, B]`` one) — or ``Absence`` when nothing is declared"""
    declared = get_type_hints(func).get("return", _NOT_DECLARED)
    if declared is _NOT_DECLARED:
        return Absence()
    return Result(declared if isinstance(declared, tuple) else (declared,))


def get_default_args(func: Callable) -> Dict[str, Any]:
    defaults: Dict[str, Any] = {}
    for name, parameter in inspect.signature(func).parameters.items():
        if parameter.default is not parameter.empty:
            defaults[name] = parameter.default
    return defaults


def check_types_serialization_compatible(annotation: Type, typ: Type, registry: SerializerRegistry) -> bool:
    """would the serializer that the registry picks for `annotation` (for any member of a Union) take a `typ`?"""
    if annotation == typ:
        return True
    for candidate in infer_real_types(annotation):
        serializer = registry.find_serializer_by_type(candidate)
        if serializer is None:
            continue
        supported = serializer.supported_types()
        if (supported == typ) if isinstance(supported, type) else supported(typ):
            return True
    return False


def is_subtype(subtype: Type, supertype: Type) -> bool:
    try:
        kind = _EMPTY_CONTAINER_KIND.get(subtype)
    except TypeError:  # an unhashable annotation object is certainly not one of the three
        kind = None
    if kind is not None:
        subtype = supertype if is_subhint(supertype, kind) else kind
    return is_subhint(subtype, supertype)
