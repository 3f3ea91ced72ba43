"""LzyCall / LazyCallWrapper (pylzy/lzy/core/call.py:38-327).

Calling an ``@op`` inside an active workflow validates the arguments against the annotations, creates
snapshot entries for arguments, results and the exception, **snapshots every local argument right
away** (``put_data`` at call time, :62-100 — later mutation of the caller's object must not leak into
the op), queues the call and hands back lazy proxies (or real values when the workflow is eager).
Outside a workflow the op is a plain function call (:207-209)."""
from __future__ import annotations

import functools
import inspect
import uuid
import weakref
from dataclasses import dataclass
from inspect import getfullargspec
from itertools import chain, zip_longest
from typing import TYPE_CHECKING, Any, Callable, Dict, List, Mapping, Optional, Sequence, Tuple

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


@dataclass(eq=False)  # identity semantics: field-wise == would touch (and materialise) proxy arguments
class LzyCall:
    workflow: "LzyWorkflow"
    signature: CallSignature
    call_env: LzyEnvironment
    description: str
    version: str
    cache: bool
    lazy_arguments: bool

    def __post_init__(self):
        self.id: str = str(uuid.uuid4())
        self.final_env: LzyEnvironment = self.workflow.owner.env.combine(self.workflow.env).combine(self.call_env)
        self.final_env.validate()
        sign, workflow = self.signature, self.workflow
        local_data_put_tasks = []

        self.__args_entry_ids: List[str] = []
        for i, arg in enumerate(sign.args):
            if is_lzy_proxy(arg):
                eid = get_proxy_entry_id(arg)
            else:
                arg_name = sign.func.arg_names[i]
                eid = workflow.snapshot.create_entry(name=self.callable_name + "." + arg_name,
                                                     typ=sign.func.input_types[arg_name]).id
                local_data_put_tasks.append(workflow.snapshot.put_data(eid, _plain(arg)))
            self.__args_entry_ids.append(eid)

        self.__kwargs_entry_ids: Dict[str, str] = {}
        for kwarg_name, kwarg in sign.kwargs.items():
            if is_lzy_proxy(kwarg):
                eid = get_proxy_entry_id(kwarg)
            else:
                eid = workflow.snapshot.create_entry(name=self.callable_name + "." + kwarg_name,
                                                     typ=sign.func.input_types[kwarg_name]).id
                local_data_put_tasks.append(workflow.snapshot.put_data(eid, _plain(kwarg)))
            self.__kwargs_entry_ids[kwarg_name] = eid

        self.__entry_ids: List[str] = []
        for i, arg_typ in enumerate(sign.func.output_types):
            self.__entry_ids.append(workflow.snapshot.create_entry(self.callable_name + f".return_{i}", arg_typ).id)
        self.__exception_id: str = workflow.snapshot.create_entry(self.callable_name + ".exception", tuple).id

        # local data goes to the snapshot here, in the constructor: it may change before the
        # dependent op actually runs
        if local_data_put_tasks:  # (a call whose arguments are all proxies does not need a trip to the event loop)
            LzyEventLoop.gather(*local_data_put_tasks)

        # every input stays pinned until the runtime has executed this call (HBM-backed snapshots
        # release a payload once nothing can ask for it any more; see FrameSnapshot)
        pin = getattr(workflow.snapshot, "pin", None)
        if pin is not None:
            for eid in list(self.__args_entry_ids) + list(self.__kwargs_entry_ids.values()):
                pin(eid)
            for eid in self.__entry_ids:  # held until the caller has its proxies (an eager workflow runs the call first)
                pin(eid)
            if self.cache:
                for eid in self.__entry_ids:
                    workflow.snapshot.persist(eid)

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
        return self.__args_entry_ids

    @property
    def kwarg_entry_ids(self) -> Mapping[str, str]:
        return self.__kwargs_entry_ids

    @property
    def entry_ids(self) -> Sequence[str]:
        return self.__entry_ids

    @property
    def exception_id(self) -> str:
        return self.__exception_id

    def get_provisioning(self) -> Provisioning:
        return self.final_env.get_provisioning()

    def get_env_vars(self) -> Dict[str, str]:
        return self.final_env.get_env_vars()


def _plain(value: Any) -> Any:
    """local data as it is stored: results of other ops sitting inside a list / tuple / dict stand for their values"""
    from ..proxy import _unwrap_tree
    return _unwrap_tree(value) if type(value) in (list, tuple, dict) else value


@dataclass(frozen=True)
class LazyCallWrapper(WithEnvironmentMixin):
    function: Callable[..., Any]
    output_types: Sequence[type]
    env: LzyEnvironment
    description: str
    version: str
    cache: bool
    lazy_arguments: bool

    def __post_init__(self):
        for attr in functools.WRAPPER_ASSIGNMENTS:
            try:
                object.__setattr__(self, attr, getattr(self.function, attr))
            except (AttributeError, TypeError):
                pass
        object.__setattr__(self, "__wrapped__", self.function)

    def with_env(self, env: LzyEnvironment) -> "LazyCallWrapper":
        return LazyCallWrapper(self.function, self.output_types, env, self.description, self.version, self.cache,
                               self.lazy_arguments)

    def __call__(self, *args, **kwargs) -> Any:
        from .workflow import LzyWorkflow

        active_workflow: Optional[LzyWorkflow] = LzyWorkflow.get_active()
        if active_workflow is None:
            return self.function(*args, **kwargs)

        signature = infer_and_validate_call_signature(
            self.function, *args, output_type=self.output_types, snapshot=active_workflow.snapshot,
            serializer_registry=active_workflow.owner.serializer_registry, **kwargs)
        lzy_call = LzyCall(workflow=active_workflow, signature=signature, call_env=self.env,
                           description=self.description, version=self.version, cache=self.cache,
                           lazy_arguments=self.lazy_arguments)
        active_workflow.register_call(lzy_call)

        single_output = len(self.output_types) == 1
        # `-> None` ops return the real None (so `is None` works) but stay queued (:232-250)
        snap = active_workflow.snapshot
        if single_output and inspect.isclass(self.output_types[0]) and issubclass(self.output_types[0], type(None)):
            if hasattr(snap, "unpin"):
                snap.unpin(lzy_call.entry_ids[0])
            return None

        proxies = []
        for eid, types in zip(lzy_call.entry_ids, lzy_call.signature.func.output_types):
            p = lzy_proxy(entry_id=eid, types=infer_real_types(types), wflow=active_workflow)
            if hasattr(snap, "pin"):  # the result's frame lives as long as this proxy can still ask for it
                snap.pin(eid)
                weakref.finalize(p, snap.unpin, eid)
            proxies.append(p)
        result = tuple(materialize(p) for p in proxies) if active_workflow.eager else tuple(proxies)
        if hasattr(snap, "unpin"):
            for eid in lzy_call.entry_ids:
                snap.unpin(eid)  # the construction-time pin; the proxies carry their own
        return result[0] if single_output else result


def infer_and_validate_call_signature(f: Callable, *args, output_type: Sequence[type], snapshot: Snapshot,
                                      serializer_registry: SerializerRegistry, **kwargs) -> CallSignature:
    types_mapping = {}
    argspec, annotations, defaults = _function_facts(f)

    if argspec.varkw is None:
        for kwarg in kwargs:
            if kwarg not in annotations:
                raise KeyError(f"Unexpected key argument {kwarg}")

    for spec_name, real_arg in zip_longest(argspec.args, range(len(args))):
        if spec_name is not None and real_arg is None and spec_name not in kwargs and spec_name not in defaults:
            raise KeyError(f"Argument {spec_name} is required but not provided")
        if argspec.varargs is None and real_arg is not None and spec_name is None:
            raise KeyError(f"Unexpected argument at position {real_arg}")

    names_for_varargs = [str(uuid.uuid4())[:8] for _ in range(len(args) - len(argspec.args))]
    actual_args: List[Any] = [None] * len(args)
    actual_kwargs = {}

    for i, (name, arg) in enumerate(chain(zip(chain(argspec.args, names_for_varargs), args), kwargs.items())):
        from_varargs = len(argspec.args) < i < len(args)
        if from_varargs:
            inferred_type = _infer_type(snapshot, name, arg)
        else:
            inferred_type = _infer_type(snapshot, name, arg, annotations)
            if name in annotations:
                typ = inferred_type if is_lzy_proxy(arg) else get_type(arg)
                compatible = check_types_serialization_compatible(annotations[name], typ, serializer_registry)
                if not compatible or not is_subtype(typ, annotations[name]):
                    raise TypeError(f"Invalid types: argument {name} has type {annotations[name]} but passed type {typ}")
        types_mapping[name] = inferred_type
        # a proxy that has already been materialised is treated as local data
        value = materialize(arg) if is_lzy_proxy(arg) and materialized(arg) else arg
        if i < len(args):
            actual_args[i] = value
        else:
            actual_kwargs[name] = value

    arg_names = tuple(argspec.args[: len(args)] + names_for_varargs)
    return CallSignature(FuncSignature(f, types_mapping, output_type, arg_names, tuple(kwargs.keys())),
                         tuple(actual_args), actual_kwargs)


_FACTS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _function_facts(f: Callable):
    """argspec, evaluated annotations and defaults of an op body, computed once per function"""
    try:
        facts = _FACTS.get(f)
    except TypeError:
        facts = None
    if facts is None:
        facts = (getfullargspec(f), inspect.get_annotations(f, eval_str=True), get_default_args(f))
        try:
            _FACTS[f] = facts
        except TypeError:
            pass
    return facts


def _infer_type(snapshot: Snapshot, arg_name: str, arg: Any, type_annotations: Optional[Dict[str, Any]] = None):
    if is_lzy_proxy(arg):
        return snapshot.get(get_proxy_entry_id(arg)).typ
    if type_annotations and arg_name in type_annotations:
        return type_annotations[arg_name]
    return get_type(arg)
