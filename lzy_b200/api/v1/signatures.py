"""What is known about an ``@op`` when it is called: the function with the types its parameters and results were
given (annotation, inference or the producing op — core/call.py), and the concrete arguments of one call.  Same
attribute surface as pylzy/lzy/api/v1/signatures.py:19-46, which the reference's tests read
(tests/core/test_calls.py)."""
from __future__ import annotations

from typing import Any, Callable, Dict, Iterator, List, Mapping, Sequence, Tuple, Type


class FuncSignature:
    __slots__ = ("callable", "input_types", "output_types", "arg_names", "kwarg_names")

    def __init__(self, callable: Callable, input_types: Mapping[str, Type], output_types: Sequence[Type],
                 arg_names: Sequence[str], kwarg_names: Sequence[str]):
        self.callable = callable
        self.input_types: Dict[str, Type] = dict(input_types)
        self.output_types: Sequence[Type] = output_types
        self.arg_names: Tuple[str, ...] = tuple(arg_names)
        self.kwarg_names: Tuple[str, ...] = tuple(kwarg_names)

    @property
    def name(self) -> str:
        return self.callable.__name__

    @property
    def param_names(self) -> List[str]:
        """every parameter that became an entry, positional ones first"""
        return [*self.input_types]

    def _key(self):
        return self.callable, self.input_types, tuple(self.output_types), self.arg_names, self.kwarg_names

    def __eq__(self, other: object) -> bool:
        return isinstance(other, FuncSignature) and self._key() == other._key()

    __hash__ = None  # type: ignore[assignment]

    def __repr__(self) -> str:
        ins = ", ".join(f"{n}: {getattr(t, '__name__', t)}" for n, t in self.input_types.items())
        outs = ", ".join(str(getattr(t, "__name__", t)) for t in self.output_types)
        return f"FuncSignature({self.name}({ins}) -> ({outs}))"


class CallSignature:
    __slots__ = ("func", "args", "kwargs")

    def __init__(self, func: FuncSignature, args: Sequence[Any], kwargs: Mapping[str, Any]):
        self.func = func
        self.args: Tuple[Any, ...] = tuple(args)
        self.kwargs: Dict[str, Any] = dict(kwargs)

    def named_arguments(self) -> Iterator[Tuple[str, Any]]:
        yield from zip(self.func.arg_names, self.args)
        yield from self.kwargs.items()

    def exec(self) -> Any:
        return self.func.callable(*self.args, **self.kwargs)

    def __repr__(self) -> str:
        return f"CallSignature({self.func.name}, {len(self.args)} args, kwargs={sorted(self.kwargs)})"
