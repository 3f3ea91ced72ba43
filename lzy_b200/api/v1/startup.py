"""The op-side entry point of a task (pylzy/lzy/api/v1/startup.py:31-224), restated over slot frames.

In the reference the worker launches ``python startup.py <pickled ProcessingRequest>``; the script resolves
every slot path against the ``LZY_MOUNT`` directory, where the slots layer materialised the inputs as files and
picks the outputs up from.  Here a slot path resolves against a :class:`SlotMount` — the worker's
``SlotsExecutionContext`` — so an argument is a zero-copy view of the HBM frame ``beforeExecution`` delivered
and a result is packed straight into the output slot's frame.  Everything else is kept: the
``ProcessingRequest`` fields, lazy arguments, the type check of results, the exception tuple written to the
exception slot and re-raised, ``None`` in the exception slot on success, results written by the main process
only (``LZY_OP_MAIN_PID``).
"""
from __future__ import annotations

import abc
import dataclasses
import logging
import logging.config
import os
import sys
import time
from logging import Logger
from typing import Any, Callable, Dict, Mapping, Optional, Sequence, Tuple, Type, cast

from ...proxy import proxy
from ...serialization.api import SerializerRegistry
from ...serialization.registry import LzySerializerRegistry, SerializerImport
from ...serialization.types import get_type
from ...utils import tbpickle
from .utils.pickle import unpickle
from .utils.types import check_types_serialization_compatible, infer_real_types, is_subtype

NAME = __name__
MAIN_PID_ENV_VAR = "LZY_OP_MAIN_PID"
_LAZY_ARGUMENT = "__lzy_lazy_argument__"


class SlotMount(abc.ABC):
    """What the ``LZY_MOUNT`` directory is to the reference's startup script (startup.py:39-41,53-54): the
    thing slot paths are resolved against."""

    @abc.abstractmethod
    def read(self, path: str, typ: Type, serializers: SerializerRegistry) -> Any:
        """the fully delivered input slot `path`, deserialized"""

    @abc.abstractmethod
    def write(self, path: str, typ: Type, data: Any, serializers: SerializerRegistry) -> None:
        """serialize `data` into the output slot `path` and mark it complete"""


class FrameMount(SlotMount):
    """Slot paths → HBM frames of one ``SlotsExecutionContext`` on one GPU."""

    def __init__(self, ctx, store, gpu: int, stream_ptr: int = 0):
        self._ctx, self._store, self._gpu, self._stream = ctx, store, gpu, stream_ptr
        self._keep = []  # packed sources stay alive until the frames are committed

    def read(self, path: str, typ: Type, serializers: SerializerRegistry) -> Any:
        from ... import _native as N
        slot = N.SlotHandle.adopt(self._ctx.input_handle(path), owned=True)
        return self._store.unpack_slot(slot, typ, serializers, private=True)

    def write(self, path: str, typ: Type, data: Any, serializers: SerializerRegistry) -> None:
        from ... import _native as N

        def factory(nbytes, device):
            return N.SlotHandle.adopt(self._ctx.output_open(path, nbytes), owned=False)

        self._keep.append(self._store.pack_object(data, typ, serializers, factory, self._stream))
        self._ctx.output_done(path)


def read_data(path: str, typ: Type, serializers: SerializerRegistry, logger: Logger, mount: SlotMount,
              cache: Optional[Dict[str, Any]] = None) -> Any:
    """one input slot as a Python object (pylzy startup.py:31-49).  Within one execution a slot is read once: the
    `cache` dict plays the part of the reference's per-process read cache."""
    if cache is not None and path in cache:
        return cache[path]
    chosen = serializers.find_serializer_by_type(typ)
    logger.info("Reading %s with serializer %s", path.rsplit("/", 1)[-1], type(chosen).__name__)
    value = mount.read(path, typ, serializers)
    if cache is not None:
        cache[path] = value
    return value


def write_data(path: str, typ: Type, data: Any, serializers: SerializerRegistry, logger: Logger, mount: SlotMount) -> None:
    """one result into its output slot (pylzy startup.py:52-77): the declared type needs a usable serializer, a lazy
    argument handed back as a result stands for its value, and the value's real type must fit the declaration."""
    chosen = serializers.find_serializer_by_type(typ)
    if chosen is None:
        raise ValueError(f"Cannot find serializer for type {typ}")
    if not chosen.available():
        raise ValueError(f"Serializer for type {typ} is not available, please install {chosen.requirements()}")
    if hasattr(type(data), _LAZY_ARGUMENT):
        data = data.__lzy_origin__
    actual = get_type(data)
    if not (check_types_serialization_compatible(typ, actual, serializers) and is_subtype(actual, typ)):
        raise TypeError(f"Invalid types: return value has type {typ} but passed type {actual}")
    logger.info("Writing %s with serializer %s", path.rsplit("/", 1)[-1], type(chosen))
    mount.write(path, typ, data, serializers)


def _get_main_pid() -> int:
    """A program that replicates itself through subprocesses (DDP) runs this entry point again in every replica.
    The first process to get here leaves its pid in ``LZY_OP_MAIN_PID`` for its children to inherit; only that
    process writes results (pylzy startup.py:80-106)."""
    return int(os.environ.setdefault(MAIN_PID_ENV_VAR, str(os.getpid())))


class _Execution:
    """one run of one op over its slots: arguments in, body, results (or the exception) out"""

    def __init__(self, serializers: SerializerRegistry, logger: Logger, mount: SlotMount, exception_path: Tuple[Type, str],
                 lazy_arguments: bool):
        self.serializers, self.logger, self.mount = serializers, logger, mount
        self.exception_type, self.exception_slot = exception_path
        self.lazy_arguments = lazy_arguments
        self._already_read: Dict[str, Any] = {}

    def put(self, typ: Type, path: str, value: Any) -> None:
        write_data(path, typ, value, self.serializers, self.logger, self.mount)

    def report_failure(self, what: str, error: BaseException, *details: Any) -> None:
        """log, then leave ``sys.exc_info()`` in the exception slot (the caller re-raises from it)"""
        self.logger.error(what, error, *details)
        self.put(self.exception_type, self.exception_slot, sys.exc_info())

    def argument(self, typ: Type, path: str) -> Any:
        def fetch() -> Any:
            return read_data(path, typ, self.serializers, self.logger, self.mount, self._already_read)

        if self.lazy_arguments:  # read on first touch, if ever
            return proxy(fetch, infer_real_types(typ), cls_attrs={_LAZY_ARGUMENT: True})
        return fetch()

    def publish(self, output_paths: Sequence[Tuple[Type, str]], result: Any) -> None:
        values = [result] if len(output_paths) == 1 else result
        for (typ, path), value in zip(output_paths, values):
            self.put(typ, path, value)
        self.put(type(None), self.exception_slot, None)  # "no exception"


def process_execution(
    serializers: SerializerRegistry,
    op: Callable,
    args_paths: Sequence[Tuple[Type, str]],
    kwargs_paths: Mapping[str, Tuple[Type, str]],
    output_paths: Sequence[Tuple[Type, str]],
    exception_path: Tuple[Type, str],
    logger: Logger,
    lazy_arguments: bool,
    mount: SlotMount,
) -> None:
    """pylzy startup.py:109-170 — the order of effects is kept: arguments are read (or wrapped lazily), the op runs,
    the main-pid marker is dropped, the main process writes every result and then ``None`` into the exception slot;
    a failure in any phase leaves ``sys.exc_info()`` in the exception slot and is re-raised."""
    run = _Execution(serializers, logger, mount, exception_path, lazy_arguments)
    me, main_pid = os.getpid(), _get_main_pid()
    logger.debug("Starting process with pid %d (ppid %d, main pid %s)", me, os.getppid(), main_pid)

    logger.info("Reading arguments...")
    try:
        positional = [run.argument(typ, path) for typ, path in args_paths]
        named = {name: run.argument(*described) for name, described in kwargs_paths.items()}
    except Exception as e:
        run.report_failure("Error while reading arguments: %s", e)
        raise

    logger.info("Executing operation '%s'", getattr(op, "__name__", op))
    began = time.time()
    try:
        result = op(*positional, **named)
    except Exception as e:
        run.report_failure("Execution completed with error `%s` in %s", e, time.time() - began)
        raise
    finally:
        os.environ.pop(MAIN_PID_ENV_VAR, None)
    logger.info("Execution completed in %s sec", time.time() - began)

    if me != main_pid:
        logger.debug("Don't write results from not main process with pid %d", me)
        return
    logger.info("Writing results...")
    try:
        run.publish(output_paths, result)
    except Exception as e:
        run.report_failure("Error while writing result: %s", e)
        raise


@dataclasses.dataclass
class ProcessingRequest:
    """what the client pickles into the task's command line (pylzy startup.py:173-182) — field for field, so that a
    request pickled by a stock client has the shape this entry point expects"""
    logger_config: Dict[str, Any]
    serializers: Sequence[SerializerImport]
    op: Callable
    args_paths: Sequence[Tuple[Type, str]]
    kwargs_paths: Mapping[str, Tuple[Type, str]]
    output_paths: Sequence[Tuple[Type, str]]
    exception_path: Tuple[Type, str]
    lazy_arguments: bool = True


def _stage(what: str, action: Callable[[], Any]) -> Any:
    """the preparatory stages report their failure on stderr (there is no logger yet) and let it propagate"""
    try:
        return action()
    except Exception as e:
        sys.stderr.write(f"Error while {what}: {e}")
        sys.stderr.flush()
        raise


def main(arg: str, mount: Optional[SlotMount] = None) -> None:
    """pylzy startup.py:185-220: unpickle the request, import the user's serializers, configure logging, run.  The
    `mount` comes from the worker that runs the task (the reference takes ``LZY_MOUNT`` from the environment and
    asserts that it is set, :39-40)."""
    if mount is None:
        raise AssertionError("no slot mount: this entry point runs inside a worker that owns the task's slots")
    request = cast(ProcessingRequest, _stage("unpickling request", lambda: unpickle(arg)))
    tbpickle.install()  # the exception slot carries sys.exc_info(), traceback included
    registry = LzySerializerRegistry()
    _stage("loading serializers", lambda: registry.load_imports(request.serializers))
    if request.logger_config:
        _stage("logging configuration", lambda: logging.config.dictConfig(request.logger_config))
    logger = logging.getLogger(NAME)
    logger.info("Starting execution...")
    process_execution(registry, request.op, request.args_paths, request.kwargs_paths, request.output_paths,
                      request.exception_path, logger, request.lazy_arguments, mount)
    logger.info("Finishing execution...")
