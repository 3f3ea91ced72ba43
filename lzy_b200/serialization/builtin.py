"""Built-in serializers: HBM-resident tensor/ndarray frames, the File format of the reference, and the
generic cloudpickle fallback.

* ``TensorFrameSerializer`` / ``NdarrayFrameSerializer`` replace serialzy's array encodings on the
  data plane (reference call sites: pylzy/lzy/api/v1/snapshot.py:138,149-150,
  pylzy/lzy/api/v1/startup.py:47,76): the payload is packed into a slot frame by the CUDA pack
  kernel; as a byte stream (legacy storage) it is ``lzyb_frame_header_t`` + raw C-order bytes.
* ``FileSerializer`` keeps the reference's byte format exactly
  (pylzy/lzy/serialization/file.py:16-50): ``b"permissions"`` + ``oct(mode & 0o777)`` (5 bytes) + raw
  bytes, and accepts header-less streams.
* ``CloudpickleSerializer`` — everything else (serialzy's encodings for those types are not
  vendored; parity is at the value level).
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import uuid
from typing import Any, BinaryIO, Callable, Dict, Optional, Tuple, Type, Union

import cloudpickle

from ..types import File
from .api import DeviceFrameSerializer, Schema, Serializer, StandardDataFormats, StandardSchemaFormats

try:  # numpy / torch are plumbing; the registry works without them
    import numpy as np
except Exception:  # pragma: no cover
    np = None  # type: ignore
try:
    import torch
except Exception:  # pragma: no cover
    torch = None  # type: ignore

HEADER_BYTES = 128
FLAG_COMMITTED, FLAG_HOST_ORIGIN, FLAG_NDARRAY, FLAG_OOB = 1, 2, 4, 8


def _np_to_code():
    from .. import _native as N
    return {"uint8": N.DT_U8, "int8": N.DT_I8, "int16": N.DT_I16, "int32": N.DT_I32, "int64": N.DT_I64,
            "float16": N.DT_F16, "float32": N.DT_F32, "float64": N.DT_F64, "bool": N.DT_BOOL,
            "uint16": N.DT_U16, "uint32": N.DT_U32, "uint64": N.DT_U64}


_CODE_TO_NP: dict = {}


def _code_to_np():
    if not _CODE_TO_NP:
        _CODE_TO_NP.update({v: k for k, v in _np_to_code().items()})
    return _CODE_TO_NP


_LOG = logging.getLogger(__name__)


class CloudpickleSerializer(Serializer):
    DATA_FORMAT = StandardDataFormats.pickle.name

    def supported_types(self) -> Union[Type, Callable[[Type], bool]]:
        return lambda t: True

    def data_format(self) -> str:
        return self.DATA_FORMAT

    def meta(self) -> Dict[str, str]:
        return {"cloudpickle": cloudpickle.__version__}

    def schema(self, typ: type) -> Schema:
        return Schema(self.DATA_FORMAT, StandardSchemaFormats.pickled_type.name, meta=self.meta())

    def _serialize(self, obj: Any, dest: BinaryIO) -> None:
        cloudpickle.dump(obj, dest)

    def _deserialize(self, source: BinaryIO, schema_type: Type, user_type: Optional[Type] = None) -> Any:
        return cloudpickle.load(source)

    # pickle protocol 5 with out-of-band buffers: large contiguous buffers inside the object (numpy
    # arrays of a Bunch / dict / dataclass, pandas blocks) are not copied into the pickle stream; they
    # travel as separate parts of the same slot frame (FrameStore.pack_object)
    @staticmethod
    def dumps_oob(obj: Any):
        buffers = []
        skeleton = cloudpickle.dumps(obj, protocol=5, buffer_callback=buffers.append)
        views = []
        for b in buffers:
            mv = b.raw()
            views.append(mv)
        return skeleton, views

    @staticmethod
    def loads_oob(skeleton, buffers) -> Any:
        import pickle
        return pickle.loads(skeleton, buffers=buffers)

    def resolve(self, schema: Schema) -> Type:
        if schema.data_format != self.DATA_FORMAT:
            raise ValueError(f"Invalid data format {schema.data_format}")
        return object


def _version_key(v: str):
    parts = []
    for piece in str(v).split("."):
        digits = "".join(ch for ch in piece if ch.isdigit())
        parts.append(int(digits) if digits else 0)
    return tuple(parts)


class FileSerializer(Serializer):
    DATA_FORMAT = "raw_file"
    PERMISSIONS_HEADER = "permissions".encode("utf-8")

    def supported_types(self) -> Union[Type, Callable[[Type], bool]]:
        return File

    def data_format(self) -> str:
        return self.DATA_FORMAT

    def stable(self) -> bool:
        return True

    def meta(self) -> Dict[str, str]:
        from .. import __version__
        return {"pylzy": __version__}

    def _serialize(self, obj: File, dest: BinaryIO) -> None:
        mode = oct(obj.stat().st_mode & 0o777)
        with obj.open("rb") as f:
            dest.write(self.PERMISSIONS_HEADER)
            dest.write(mode.encode("utf-8"))
            while True:
                data = f.read(4096)
                if not data:
                    break
                dest.write(data)

    def _deserialize(self, source: BinaryIO, schema_type: Type, user_type: Optional[Type] = None) -> Any:
        self._check_types_valid(schema_type, user_type)
        new_path = os.path.join("/tmp", str(uuid.uuid1()))
        permissions: Optional[str] = None
        with open(new_path, "wb") as f:
            header = source.read(len(self.PERMISSIONS_HEADER))
            if header == self.PERMISSIONS_HEADER:
                permissions = source.read(5).decode("utf-8")
            elif len(header) > 0:
                f.write(header)  # compatibility with streams written without the header
            while True:
                data = source.read(4096)
                if not data:
                    break
                f.write(data)
        file = File(new_path)
        if permissions:
            file.chmod(int(permissions, 8))
        return file

    def schema(self, typ: type) -> Schema:
        if typ != File:
            raise ValueError(f"Invalid type {typ}")
        return Schema(self.DATA_FORMAT, StandardSchemaFormats.no_schema.name, meta=self.meta())

    def resolve(self, schema: Schema) -> Type:
        if schema.data_format != self.DATA_FORMAT:
            raise ValueError(f"Invalid data format {schema.data_format}")
        if schema.schema_format != StandardSchemaFormats.no_schema.name:
            raise ValueError(f"Invalid schema format {schema.schema_format}")
        # version drift is reported, not refused (pylzy/lzy/serialization/file.py:72-76)
        from .. import __version__
        if "pylzy" not in schema.meta:
            _LOG.warning("No pylzy version in meta")
        elif _version_key(schema.meta["pylzy"]) > _version_key(__version__):
            _LOG.warning("Installed version of pylzy %s is older than used for serialization %s", __version__,
                         schema.meta["pylzy"])
        return File


def _header_bytes(dtype_code: int, shape, nbytes: int, flags: int) -> bytes:
    from .. import _native as N
    h = N.FrameHeader()
    for i, b in enumerate(b"LZYB200\0"):
        h.magic[i] = b
    h.version, h.dtype, h.nbytes, h.ndim, h.flags = 1, dtype_code, nbytes, len(shape), flags
    for i, s in enumerate(shape):
        h.shape[i] = int(s)
    return bytes(h)


def _parse_header(raw: bytes):
    from .. import _native as N
    if len(raw) != HEADER_BYTES or raw[:8] != b"LZYB200\0":
        raise ValueError("stream is not an lzy_b200 tensor frame")
    return N.FrameHeader.from_buffer_copy(raw)


class NdarrayFrameSerializer(DeviceFrameSerializer):
    DATA_FORMAT = "lzyb_ndarray_frame"

    def available(self) -> bool:
        return np is not None

    def supported_types(self) -> Union[Type, Callable[[Type], bool]]:
        return np.ndarray if np is not None else (lambda t: False)

    def data_format(self) -> str:
        return self.DATA_FORMAT

    def stable(self) -> bool:
        return True

    def handles(self, typ: Type) -> bool:
        return np is not None and isinstance(typ, type) and issubclass(typ, np.ndarray)

    @staticmethod
    def _plan(a) -> Tuple[int, Tuple[int, ...], int]:
        code = _np_to_code().get(a.dtype.name)
        if code is None or a.ndim > 8:
            raise TypeError(f"ndarray of dtype {a.dtype} / ndim {a.ndim} cannot travel in a tensor frame")
        return code, tuple(a.shape), a.nbytes

    def frame_plan(self, obj):
        code, shape, nbytes = self._plan(obj)
        return code, shape, nbytes, FLAG_HOST_ORIGIN | FLAG_NDARRAY

    def can_frame(self, obj) -> bool:
        return np is not None and isinstance(obj, np.ndarray) and obj.dtype.name in _np_to_code() and obj.ndim <= 8

    def pack(self, obj, slot, stream: int) -> None:
        a = np.ascontiguousarray(obj)
        slot.pack_host(a.ctypes.data, a.nbytes, stream)
        slot._keepalive = a  # the H2D copy is asynchronous until commit()

    def unpack(self, slot, info, private: bool):
        from .. import _native as N
        shape = tuple(int(info.shape[i]) for i in range(info.ndim))
        dtype = np.dtype(_code_to_np()[int(info.dtype)])
        nbytes = int(info.nbytes)
        # a pooled pinned block as the array's memory: the D2H lands in it directly (no staging copy, no first-touch faults)
        buf = N.pinned_buffer(nbytes) if nbytes else None
        out = np.frombuffer(buf, dtype=dtype).reshape(shape) if buf is not None else np.empty(shape, dtype=dtype)
        if nbytes:
            slot.copy_to_host(out.ctypes.data, nbytes)
        return out

    def _serialize(self, obj, dest: BinaryIO) -> None:
        a = np.require(obj, requirements="C")  # not ascontiguousarray: that turns a 0-d array into shape (1,)
        code, shape, nbytes = self._plan(a)
        dest.write(_header_bytes(code, shape, nbytes, FLAG_HOST_ORIGIN | FLAG_NDARRAY))
        dest.write(a.tobytes() if a.nbytes < (1 << 20) else memoryview(a).cast("B"))

    def _deserialize(self, source: BinaryIO, schema_type: Type, user_type: Optional[Type] = None):
        h = _parse_header(source.read(HEADER_BYTES))
        names = {v: k for k, v in _np_to_code().items()}
        shape = tuple(int(h.shape[i]) for i in range(h.ndim))
        data = source.read(int(h.nbytes))
        return np.frombuffer(data, dtype=names[int(h.dtype)]).reshape(shape).copy()


class TensorFrameSerializer(DeviceFrameSerializer):
    DATA_FORMAT = "lzyb_tensor_frame"

    def available(self) -> bool:
        return torch is not None

    def supported_types(self) -> Union[Type, Callable[[Type], bool]]:
        return torch.Tensor if torch is not None else (lambda t: False)

    def data_format(self) -> str:
        return self.DATA_FORMAT

    def stable(self) -> bool:
        return True

    def handles(self, typ: Type) -> bool:
        return torch is not None and isinstance(typ, type) and issubclass(typ, torch.Tensor)

    def can_frame(self, obj) -> bool:
        if torch is None or not isinstance(obj, torch.Tensor) or obj.dim() > 8 or obj.is_sparse or obj.is_quantized:
            return False
        from ..devmem import _TORCH_TO_DT
        return obj.dtype in _TORCH_TO_DT and not obj.requires_grad

    def frame_plan(self, obj):
        from ..devmem import dtype_code
        flags = 0 if obj.is_cuda else FLAG_HOST_ORIGIN
        return dtype_code(obj.dtype), tuple(obj.shape), obj.numel() * obj.element_size(), flags

    def pack(self, obj, slot, stream: int) -> None:
        nbytes = obj.numel() * obj.element_size()
        if not obj.is_cuda:
            t = obj.contiguous()
            slot.pack_host(t.data_ptr(), nbytes, stream)
            slot._keepalive = t
            return
        if obj.device.index != slot.gpu:
            obj = obj.to(f"cuda:{slot.gpu}")
        if obj.is_contiguous():
            slot.pack(obj.data_ptr(), nbytes, stream)
        elif obj.dim() == 2 and obj.stride(1) == 1 and (obj.shape[1] * obj.element_size()) % 16 == 0:
            # a row-pitched view (column window of a matrix): the gather kernel packs it in one pass
            slot.pack_2d(obj.data_ptr(), obj.shape[0], obj.shape[1] * obj.element_size(), obj.stride(0) * obj.element_size(), stream)
        else:
            t = obj.contiguous()
            slot.pack(t.data_ptr(), nbytes, stream)
            slot._keepalive = t
            return
        slot._keepalive = obj

    def unpack(self, slot, info, private: bool):
        from ..devmem import torch_dtype, view_tensor
        shape = tuple(int(info.shape[i]) for i in range(info.ndim))
        dt = torch_dtype(int(info.dtype))
        host_origin = bool(getattr(info, "flags", 0) & FLAG_HOST_ORIGIN)
        if host_origin:
            out = torch.empty(shape, dtype=dt)
            if out.numel():
                slot.copy_to_host(out.data_ptr(), out.numel() * out.element_size())
            return out
        # zero-copy view of the frame; the holder keeps the frame alive as long as the tensor lives
        return view_tensor(slot.device_ptr(), dt, shape, slot.gpu, owner=slot)

    def _serialize(self, obj, dest: BinaryIO) -> None:
        code, shape, nbytes, flags = self.frame_plan(obj)
        dest.write(_header_bytes(code, shape, nbytes, flags))
        t = obj.detach().contiguous().cpu()
        if nbytes:
            dest.write((C.c_char * nbytes).from_address(t.data_ptr()).raw)

    def _deserialize(self, source: BinaryIO, schema_type: Type, user_type: Optional[Type] = None):
        from ..devmem import torch_dtype
        h = _parse_header(source.read(HEADER_BYTES))
        shape = tuple(int(h.shape[i]) for i in range(h.ndim))
        out = torch.empty(shape, dtype=torch_dtype(int(h.dtype)))
        data = source.read(int(h.nbytes))
        if data:
            C.memmove(out.data_ptr(), data, len(data))
        return out if (h.flags & FLAG_HOST_ORIGIN) or not torch.cuda.is_available() else out.cuda()
