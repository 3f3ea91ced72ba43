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
    """startup.py:31-49; the read cache lives for one execution (the reference's is per process = per op)."""
    if cache is not None and path in cache:
        return cache[path]
    ser = serializers.find_serializer_by_type(typ)
    logger.info("Reading %s with serializer %s", path.split("/")[-1], type(ser).__name__)
    data = mount.read(path, typ, serializers)
    if cache is not None:
        cache[path] = data
    return data


def write_data(path: str, typ: Type, data: Any, serializers: SerializerRegistry, logger: Logger, mount: SlotMount) -> None:
    """startup.py:52-77: serializer must exist and be available; a lazy argument returned as a result is
    unwrapped; the value's real type must be compatible with the annotation."""
    ser = serializers.find_serializer_by_type(typ)
    if ser is None:
        raise ValueError(f"Cannot find serializer for type {typ}")
    if not ser.available():
        raise ValueError(f"Serializer for type {typ} is not available, please install {ser.requirements()}")
    if hasattr(type(data), _LAZY_ARGUMENT):  # if input argument is a return value
        data = data.__lzy_origin__
    real_type = get_type(data)
    compatible = check_types_serialization_compatible(typ, real_type, serializers)
    if not compatible or not is_subtype(real_type, typ):
        raise TypeError(f"Invalid types: return value has type {typ} but passed type {real_type}")
    logger.info("Writing %s with serializer %s", path.split("/")[-1], type(ser))
    mount.write(path, typ, data, serializers)


def _get_main_pid() -> int:
    """startup.py:80-106: a program replicated by subprocesses (DDP) re-runs this entry point; only the
    initial process — the one that set LZY_OP_MAIN_PID — writes results."""
    if MAIN_PID_ENV_VAR not in os.environ:
        os.environ[MAIN_PID_ENV_VAR] = str(os.getpid())
    return int(os.environ[MAIN_PID_ENV_VAR])


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
    """startup.py:109-170, step for step."""
    pid = os.getpid()
    main_pid = _get_main_pid()
    logger.debug("Starting process with pid %d (ppid %d, main pid %s)", pid, os.getppid(), main_pid)
    logger.info("Reading arguments...")
    exc_typ, exc_path = exception_path
    cache: Dict[str, Any] = {}

    def argument(typ: Type, path: str) -> Any:
        if lazy_arguments:
            return proxy(lambda: read_data(path, typ, serializers, logger, mount, cache), infer_real_types(typ),
                         cls_attrs={_LAZY_ARGUMENT: True})
        return read_data(path, typ, serializers, logger, mount, cache)

    try:
        args = [argument(typ, path) for typ, path in args_paths]
        kwargs = {name: argument(typ, path) for name, (typ, path) in kwargs_paths.items()}
    except Exception as e:
        logger.error("Error while reading arguments: %s", e)
        write_data(exc_path, exc_typ, sys.exc_info(), serializers, logger, mount)
        raise

    logger.info("Executing operation '%s'", getattr(op, "__name__", op))
    start = time.time()
    try:
        res = op(*args, **kwargs)
    except Exception as e:
        logger.error("Execution completed with error `%s` in %s", e, time.time() - start)
        write_data(exc_path, exc_typ, sys.exc_info(), serializers, logger, mount)
        raise
    finally:
        os.environ.pop(MAIN_PID_ENV_VAR, None)
    logger.info("Execution completed in %s sec", time.time() - start)

    if main_pid != pid:
        logger.debug("Don't write results from not main process with pid %d", pid)
        return

    logger.info("Writing results...")
    try:
        if len(output_paths) == 1:
            write_data(output_paths[0][1], output_paths[0][0], res, serializers, logger, mount)
        else:
            for out, data in zip(output_paths, res):
                write_data(out[1], out[0], data, serializers, logger, mount)
        write_data(exc_path, type(None), None, serializers, logger, mount)
    except Exception as e:
        logger.error("Error while writing result: %s", e)
        write_data(exc_path, exc_typ, sys.exc_info(), serializers, logger, mount)
        raise


@dataclasses.dataclass
class ProcessingRequest:
    """startup.py:173-182 — field for field, so a request pickled by a stock client has the same shape."""
    logger_config: Dict[str, Any]
    serializers: Sequence[SerializerImport]
    op: Callable
    args_paths: Sequence[Tuple[Type, str]]
    kwargs_paths: Mapping[str, Tuple[Type, str]]
    output_paths: Sequence[Tuple[Type, str]]
    exception_path: Tuple[Type, str]
    lazy_arguments: bool = True


def main(arg: str, mount: Optional[SlotMount] = None) -> None:
    """startup.py:185-220.  `mount` is supplied by the worker that runs the task (the reference reads
    ``LZY_MOUNT`` from the environment and asserts it is set, startup.py:39-40)."""
    if mount is None:
        raise AssertionError("no slot mount: this entry point runs inside a worker that owns the task's slots")
    try:
        req = cast(ProcessingRequest, unpickle(arg))
    except Exception as e:
        sys.stderr.write(f"Error while unpickling request: {e}")
        sys.stderr.flush()
        raise
    tbpickle.install()  # the exception slot carries sys.exc_info(), traceback included
    registry = LzySerializerRegistry()
    try:
        registry.load_imports(req.serializers)
    except Exception as e:
        sys.stderr.write(f"Error while loading serializers: {e}")
        sys.stderr.flush()
        raise
    try:
        if req.logger_config:
            logging.config.dictConfig(req.logger_config)
    except Exception as e:
        sys.stderr.write(f"Error while logging configuration: {e}")
        sys.stderr.flush()
        raise
    logger = logging.getLogger(NAME)
    logger.info("Starting execution...")
    process_execution(registry, req.op, req.args_paths, req.kwargs_paths, req.output_paths, req.exception_path, logger,
                      req.lazy_arguments, mount)
    logger.info("Finishing execution...")
