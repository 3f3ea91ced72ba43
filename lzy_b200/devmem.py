"""Zero-copy views of slot payloads as torch tensors (PyTorch is plumbing here: device memory,
streams).  A slot's HBM frame is owned by liblzy_b200; the view must not outlive the slot."""
from __future__ import annotations

import contextlib
from typing import Sequence

import torch

from . import _native as N
from .utils.event_loop import caller_stream

_TORCH_TO_DT = {
    torch.uint8: N.DT_U8, torch.int8: N.DT_I8, torch.int16: N.DT_I16, torch.int32: N.DT_I32,
    torch.int64: N.DT_I64, torch.float16: N.DT_F16, torch.bfloat16: N.DT_BF16,
    torch.float32: N.DT_F32, torch.float64: N.DT_F64, torch.bool: N.DT_BOOL,
}
for _name, _code in (("uint16", N.DT_U16), ("uint32", N.DT_U32), ("uint64", N.DT_U64)):
    if hasattr(torch, _name):
        _TORCH_TO_DT[getattr(torch, _name)] = _code
_DT_TO_TORCH = {v: k for k, v in _TORCH_TO_DT.items()}
_DT_TO_TORCH[N.DT_BYTES] = torch.uint8
_DT_TO_TORCH[N.DT_PICKLE] = torch.uint8


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _TORCH_TO_DT[dt]
    except KeyError:
        raise TypeError(f"dtype {dt} cannot travel in a tensor frame") from None


def torch_dtype(code: int) -> torch.dtype:
    return _DT_TO_TORCH[code]


class _RawCuda:
    """Minimal __cuda_array_interface__ carrier for a raw device pointer."""

    def __init__(self, ptr: int, nbytes: int, owner=None):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None,
        }
        self._owner = owner


def view_bytes(ptr: int, nbytes: int, gpu: int, owner=None) -> torch.Tensor:
    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device=f"cuda:{gpu}")
    with torch.cuda.device(gpu):
        return torch.as_tensor(_RawCuda(ptr, nbytes, owner), device=f"cuda:{gpu}")


def view_tensor(ptr: int, dtype: torch.dtype, shape: Sequence[int], gpu: int, owner=None) -> torch.Tensor:
    n = 1
    for s in shape:
        n *= int(s)
    nbytes = n * torch.empty(0, dtype=dtype).element_size()
    return view_bytes(ptr, nbytes, gpu, owner).view(dtype).view(*[int(s) for s in shape])


def current_stream_ptr(gpu: int) -> int:
    """cudaStream_t of torch's current stream.  The C ABI reserves NULL for "the library's own
    stream", so the legacy default stream (handle 0) is passed as cudaStreamLegacy (0x1)."""
    return int(torch.cuda.current_stream(gpu).cuda_stream) or 1


@contextlib.contextmanager
def caller_stream_scope(gpu: int):
    """Inside a coroutine started with ``LzyEventLoop.run_async``: make the *caller's* current stream on `gpu`
    the current one for the block, so that work on a tensor the caller handed in (or will receive) is ordered
    with the caller's own work on it.  Elsewhere: a no-op."""
    s = caller_stream(gpu)
    if s is None:
        yield
        return
    with torch.cuda.stream(s):
        yield
