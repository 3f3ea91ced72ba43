"""Calling an ``@op`` — the behaviour of pylzy/lzy/core/call.py:38-327, built from three pieces:

``LazyCallWrapper``   what ``@op`` puts in place of the function.  Outside an active workflow it simply calls the
                      function (:207-209).  Inside one it binds the arguments (``infer_and_validate_call_signature``),
                      records an ``LzyCall``, queues it and answers with lazy proxies — real values when the workflow is
                      eager, the real ``None`` for ``-> None`` ops (:232-250).
``LzyCall``           one recorded call: snapshot entries for every argument, every result and the exception, the
                      environment frozen at call time (:58-61).  **Local arguments are written to the snapshot right
                      here** (:96-100): what the op will see is what the caller passed, whatever happens to the
                      caller's object afterwards.
``infer_and_validate_call_signature``  arity checks (``KeyError``), types of the arguments — from the producing op for
                      a proxy, from the annotation, else from the value — and the check of each annotated argument
                      against its annotation (``TypeError``, :303-312).

An HBM-backed snapshot (``FrameSnapshot``) additionally wants to know who can still ask for an entry: inputs stay
pinned until the runtime has run the call, results while their proxies are alive."""
from __future__ import annotations

import functools
import inspect
import uuid
import weakref
from typing import TYPE_CHECKING, Any, Callable, Dict, List, Mapping, NamedTuple, Optional, Sequence, Tuple

from ..api.v1.signatures import CallSignature, FuncSignature
from ..api.v1.snapshot import Snapshot
from ..api.v1.utils.proxy_adapter import get_proxy_entry_id, is_lzy_proxy, lzy_proxy, materialize, materialized
from ..api.v1.utils.types import check_types_serialization_compatible, get_default_args, infer_real_types, is_subtype
from ..env import LzyEnvironment, Provisioning, WithEnvironmentMixin
from ..serialization.api import SerializerRegistry
from ..serialization.types import get_type
from ..utils.event_loop import LzyEventLoop

if TYPE_CHECKING:
    from .workflow import LzyWorkflow


# ---- one recorded call ---------------------------------------------------------------------------------------------
class LzyCall:
    """identity semantics on purpose: comparing two calls field by field would compare — and so materialise — their
    proxy arguments"""

    def __init__(self, workflow: "LzyWorkflow", signature: CallSignature, call_env: LzyEnvironment, description: str,
                 version: str, cache: bool, lazy_arguments: bool):
        self.workflow, self.signature, self.call_env = workflow, signature, call_env
        self.description, self.version, self.cache, self.lazy_arguments = description, version, cache, lazy_arguments
        self.id: str = str(uuid.uuid4())
        # the environment is fixed now: later changes of the Lzy / workflow environment do not reach this call
        self.final_env: LzyEnvironment = workflow.owner.env.combine(workflow.env).combine(call_env)
        self.final_env.validate()

        snapshot, func = workflow.snapshot, signature.func
        label = func.callable.__name__
        uploads = []

        def input_entry(parameter: str, value: Any) -> str:
            if is_lzy_proxy(value):  # the result of another op: its entry already exists
                return get_proxy_entry_id(value)
            entry = snapshot.create_entry(name=f"{label}.{parameter}", typ=func.input_types[parameter])
            uploads.append(snapshot.put_data(entry.id, _plain(value)))
            return entry.id

        self._arg_entries: List[str] = [input_entry(func.arg_names[i], v) for i, v in enumerate(signature.args)]
        self._kwarg_entries: Dict[str, str] = {name: input_entry(name, v) for name, v in signature.kwargs.items()}
        self._result_entries: List[str] = [snapshot.create_entry(f"{label}.return_{i}", typ).id
                                           for i, typ in enumerate(func.output_types)]
        self._exception_entry: str = snapshot.create_entry(f"{label}.exception", tuple).id
        if uploads:  # (a call whose arguments are all proxies needs no trip to the workflow thread)
            LzyEventLoop.gather(*uploads, inline=bool(getattr(snapshot, "inline_coroutines", False)))

        pin = getattr(snapshot, "pin", None)
        if pin is not None:
            for eid in self.input_entry_ids():
                pin(eid)  # until the runtime has executed (or dropped) this call
            for eid in self._result_entries:
                pin(eid)  # until the caller holds its proxies (an eager workflow runs the call before that)
                if cache:
                    snapshot.persist(eid)

    def input_entry_ids(self) -> List[str]:
        return [*self._arg_entries, *self._kwarg_entries.values()]

    @property
    def callable_name(self) -> str:
        return self.signature.func.callable.__name__

    @property
    def args(self) -> Tuple[Any, ...]:
        return self.signature.args

    @property
    def kwargs(self) -> Dict[str, Any]:
        return self.signature.kwargs

    @property
    def arg_entry_ids(self) -> Sequence[str]:
        return self._arg_entries

    @property
    def kwarg_entry_ids(self) -> Mapping[str, str]:
        return self._kwarg_entries

    @property
    def entry_ids(self) -> Sequence[str]:
        return self._result_entries

    @property
    def exception_id(self) -> str:
        return self._exception_entry

    def get_provisioning(self) -> Provisioning:
        return self.final_env.get_provisioning()

    def get_env_vars(self) -> Dict[str, str]:
        return self.final_env.get_env_vars()

    def __repr__(self) -> str:
        return f"LzyCall({self.callable_name}, id={self.id})"


def _plain(value: Any) -> Any:
    """local data as it is stored: results of other ops sitting inside a list / tuple / dict stand for their values"""
    from ..proxy import _unwrap_tree
    return _unwrap_tree(value) if type(value) in (list, tuple, dict) else value


# ---- what @op leaves in place of the function ----------------------------------------------------------------------
class LazyCallWrapper(WithEnvironmentMixin):
    def __init__(self, function: Callable[..., Any], output_types: Sequence[type], env: LzyEnvironment, description: str,
                 version: str, cache: bool, lazy_arguments: bool):
        self.function, self.output_types, self.env = function, output_types, env
        self.description, self.version, self.cache, self.lazy_arguments = description, version, cache, lazy_arguments
        functools.update_wrapper(self, function)  # __name__, __doc__, __module__, __wrapped__ …

    def with_env(self, env: LzyEnvironment) -> "LazyCallWrapper":
        return LazyCallWrapper(self.function, self.output_types, env, self.description, self.version, self.cache,
                               self.lazy_arguments)

    def _returns_none(self) -> bool:
        only = self.output_types[0] if len(self.output_types) == 1 else None
        return inspect.isclass(only) and issubclass(only, type(None))

    def __call__(self, *args: Any, **kwargs: Any) -> Any:
        from .workflow import LzyWorkflow

        workflow = LzyWorkflow.get_active()
        if workflow is None:
            return self.function(*args, **kwargs)

        snapshot = workflow.snapshot
        signature = infer_and_validate_call_signature(self.function, *args, output_type=self.output_types, snapshot=snapshot,
                                                      serializer_registry=workflow.owner.serializer_registry, **kwargs)
        call = LzyCall(workflow, signature, self.env, self.description, self.version, self.cache, self.lazy_arguments)
        workflow.register_call(call)  # an eager workflow executes it right here

        pin, unpin = getattr(snapshot, "pin", None), getattr(snapshot, "unpin", None)
        answer: Any
        if self._returns_none():
            answer = None  # the real None, so that `is None` holds; the call stays queued all the same
        else:
            handles = []
            for eid, declared in zip(call.entry_ids, signature.func.output_types):
                handle = lzy_proxy(entry_id=eid, types=infer_real_types(declared), wflow=workflow)
                if pin is not None:  # the result's frame lives as long as this proxy can still ask for it
                    pin(eid)
                    weakref.finalize(handle, unpin, eid)
                handles.append(materialize(handle) if workflow.eager else handle)
            answer = handles[0] if len(handles) == 1 else tuple(handles)
        if unpin is not None:
            for eid in call.entry_ids:
                unpin(eid)  # the construction-time pin; the proxies carry their own
        return answer


# ---- binding the arguments of one call -------------------------------------------------------------------------
class _Facts(NamedTuple):
    spec: inspect.FullArgSpec
    annotations: Dict[str, Any]
    defaults: Dict[str, Any]


_FACTS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _function_facts(f: Callable) -> _Facts:
    """argspec, evaluated annotations and defaults of an op body — computed once per function"""
    try:
        known = _FACTS.get(f)
    except TypeError:  # not weak-referenceable
        known = None
    if known is None:
        known = _Facts(inspect.getfullargspec(f), inspect.get_annotations(f, eval_str=True), get_default_args(f))
        try:
            _FACTS[f] = known
        except TypeError:
            pass
    return known


def _refuse_bad_arity(facts: _Facts, n_positional: int, kwargs: Mapping[str, Any]) -> None:
    spec = facts.spec
    if spec.varkw is None:
        for name in kwargs:
            if name not in facts.annotations:
                raise KeyError(f"Unexpected key argument {name}")
    for name in spec.args[n_positional:]:
        if name not in kwargs and name not in facts.defaults:
            raise KeyError(f"Argument {name} is required but not provided")
    if spec.varargs is None and n_positional > len(spec.args):
        raise KeyError(f"Unexpected argument at position {len(spec.args)}")


def _as_local_data(value: Any) -> Any:
    """a proxy that has been materialised already is plain local data again"""
    return materialize(value) if is_lzy_proxy(value) and materialized(value) else value


def infer_and_validate_call_signature(f: Callable, *args: Any, output_type: Sequence[type], snapshot: Snapshot,
                                      serializer_registry: SerializerRegistry, **kwargs: Any) -> CallSignature:
    facts = _function_facts(f)
    _refuse_bad_arity(facts, len(args), kwargs)
    declared = facts.spec.args
    vararg_names = [uuid.uuid4().hex[:8] for _ in args[len(declared):]]  # *args become entries with made-up names
    positional_names = [*declared[:len(args)], *vararg_names]

    def type_of(name: str, value: Any, annotated: bool) -> Any:
        if is_lzy_proxy(value):
            produced_as = snapshot.get(get_proxy_entry_id(value)).typ
            actual = produced_as
        else:
            produced_as, actual = None, get_type(value)
        if not annotated:
            return actual
        wanted = facts.annotations[name]
        if not (check_types_serialization_compatible(wanted, actual, serializer_registry) and is_subtype(actual, wanted)):
            raise TypeError(f"Invalid types: argument {name} has type {wanted} but passed type {actual}")
        return produced_as if produced_as is not None else wanted

    types: Dict[str, Any] = {}
    for index, (name, value) in enumerate(zip(positional_names, args)):
        types[name] = type_of(name, value, annotated=index < len(declared) and name in facts.annotations)
    for name, value in kwargs.items():
        types[name] = type_of(name, value, annotated=name in facts.annotations)

    return CallSignature(FuncSignature(f, types, output_type, tuple(positional_names), tuple(kwargs)),
                         tuple(_as_local_data(v) for v in args), {k: _as_local_data(v) for k, v in kwargs.items()})
