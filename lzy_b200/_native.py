"""ctypes binding of liblzy_b200.so — the C ABI declared in include/lzy_b200.h.

This is the only way Python reaches the data plane.  There is no Python or CPU fallback: if the
shared library has not been built, importing this module raises; if no CUDA device is present,
``init()`` raises ``NoGpuError`` and every data call raises ``LzybError``.

The same entry points are what a JNI shim would forward to from
``ai.lzy.slots.backends.{InputSlotBackend,OutputSlotBackend}`` and
``ai.lzy.slots.transfers.InputTransfer``.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from pathlib import Path
from typing import Optional, Sequence, Tuple

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "liblzy_b200.so"

# status codes (include/lzy_b200.h)
OK, EOF = 0, -1
E_PEER, E_LOCAL, E_ARG, E_STATE, E_NOGPU = -2, -3, -4, -5, -6
E_NOT_FOUND, E_PRECONDITION, E_UNAVAILABLE, E_INTERNAL = -7, -8, -9, -10
E_UNIMPLEMENTED, E_INVALID, E_EXISTS = -11, -12, -13

ST_EMPTY, ST_FILLING, ST_READY, ST_FAILED, ST_CLOSED, ST_MOVED = 0, 1, 2, 3, 4, 5

DT_BYTES, DT_U8, DT_I8, DT_I16, DT_I32, DT_I64 = 0, 1, 2, 3, 4, 5
DT_F16, DT_BF16, DT_F32, DT_F64, DT_BOOL = 6, 7, 8, 9, 10
DT_U16, DT_U32, DT_U64, DT_PICKLE = 11, 12, 13, 32

FAULT_NONE, FAULT_PEER_AFTER, FAULT_CORRUPT_ONCE, FAULT_LOCAL_OPEN, FAULT_SLAB_RESERVE = 0, 1, 2, 3, 4
ENGINE_AUTO, ENGINE_LDG, ENGINE_TMA, ENGINE_DMA, ENGINE_PUSH = 0, 1, 2, 3, 4
MAX_DIMS = 8


class LzybError(RuntimeError):
    """Any negative status from the C ABI.  ``code`` is the LZYB_E_* value."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


class ReadException(LzybError):
    """LZYB_E_PEER — image of ``InputTransfer.ReadException``
    (lzy/slots/src/main/java/ai/lzy/slots/transfers/InputTransfer.java:13-17): the *peer* failed;
    report TransferFailed, get another peer, resume at the delivered offset."""


class NoGpuError(LzybError):
    """LZYB_E_NOGPU — the product path never runs on the CPU."""


class FrameHeader(C.Structure):
    _fields_ = [
        ("magic", C.c_uint8 * 8),
        ("version", C.c_uint32),
        ("dtype", C.c_uint32),
        ("nbytes", C.c_uint64),
        ("ndim", C.c_uint32),
        ("flags", C.c_uint32),
        ("shape", C.c_uint64 * MAX_DIMS),
        ("digest", C.c_uint64 * 2),
        ("sum", C.c_uint64),
        ("xr", C.c_uint64),
    ]


class SlotInfo(C.Structure):
    _fields_ = [
        ("nbytes", C.c_uint64),
        ("gpu", C.c_int32),
        ("state", C.c_int32),
        ("is_input", C.c_int32),
        ("is_remote", C.c_int32),
        ("delivered", C.c_uint64),
        ("digest", C.c_uint64 * 2),
        ("dtype", C.c_uint32),
        ("ndim", C.c_uint32),
        ("shape", C.c_uint64 * MAX_DIMS),
    ]


class IpcDesc(C.Structure):
    _fields_ = [
        ("ipc_handle", C.c_uint8 * 64),
        ("buffer_bytes", C.c_uint64),
        ("nbytes", C.c_uint64),
        ("gpu", C.c_int32),
        ("pid", C.c_int32),
        ("digest", C.c_uint64 * 2),
        ("sum", C.c_uint64),
        ("xr", C.c_uint64),
        ("dtype", C.c_uint32),
        ("ndim", C.c_uint32),
        ("shape", C.c_uint64 * MAX_DIMS),
        ("reserved", C.c_uint8 * 64),
    ]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "kernel_launches", "pack_launches", "pull_launches", "scan_launches", "frame_launches",
        "bytes_packed", "bytes_pulled_local", "bytes_pulled_peer", "bytes_framed_host",
        "h2d_bytes", "d2h_bytes", "slots_live", "hbm_bytes_live", "hbm_bytes_pooled", "dma_copies",
        "push_launches", "frames_moved")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


assert C.sizeof(FrameHeader) == 128 and C.sizeof(IpcDesc) == 256

if not _LIB_PATH.exists():
    raise ImportError(
        f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        f"or `make -C lzy_b200/csrc`.  lzy_b200 has no pure-Python data path.")

lib = C.CDLL(str(_LIB_PATH), mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)

_H = C.c_uint64
_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "lzyb_init": (C.c_int, [C.c_int, C.c_int, C.c_uint64]),
    "lzyb_shutdown": (C.c_int, []),
    "lzyb_is_initialized": (C.c_int, []),
    "lzyb_abi_version": (C.c_int, []),
    "lzyb_device_count": (C.c_int, []),
    "lzyb_last_error": (C.c_char_p, []),
    "lzyb_strerror": (C.c_char_p, [C.c_int]),
    "lzyb_set_last_error": (C.c_int, [C.c_char_p]),
    "lzyb_slot_create": (C.c_int, [C.c_char_p, C.c_int, C.c_uint64, C.POINTER(_H)]),
    "lzyb_slot_device_ptr": (C.c_int, [_H, C.POINTER(_P)]),
    "lzyb_slot_set_meta": (C.c_int, [_H, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
    "lzyb_slot_pack": (C.c_int, [_H, _P, C.c_uint64, _P]),
    "lzyb_slot_pack_2d": (C.c_int, [_H, _P, C.c_uint64, C.c_uint64, C.c_uint64, _P]),
    "lzyb_slot_pack_host": (C.c_int, [_H, _P, C.c_uint64, _P]),
    "lzyb_slot_pack_host_part": (C.c_int, [_H, C.c_uint64, _P, C.c_uint64, C.c_int, _P]),
    "lzyb_slot_commit": (C.c_int, [_H, C.POINTER(C.c_uint64 * 2), _P]),
    "lzyb_slot_read": (C.c_int64, [_H, C.c_uint64, _P, C.c_uint64]),
    "lzyb_slot_info": (C.c_int, [_H, C.POINTER(SlotInfo)]),
    "lzyb_slot_meta": (C.c_int, [_H, C.POINTER(SlotInfo)]),
    "lzyb_slot_wait": (C.c_int, [_H, _P]),
    "lzyb_set_lazy_commit": (C.c_int, [C.c_int]),
    "lzyb_set_transfer_waves": (C.c_int, [C.c_int]),
    "lzyb_host_alloc": (C.c_int, [C.c_uint64, C.POINTER(C.c_void_p)]),
    "lzyb_host_free": (C.c_int, [C.c_void_p]),
    "lzyb_set_host_pool_limit": (C.c_int, [C.c_uint64]),
    "lzyb_slot_header": (C.c_int, [_H, C.POINTER(FrameHeader)]),
    "lzyb_slot_close": (C.c_int, [_H]),
    "lzyb_slot_retain": (C.c_int, [_H]),
    "lzyb_slot_set_movable": (C.c_int, [_H, C.c_int]),
    "lzyb_slot_find": (C.c_int, [C.c_char_p, C.POINTER(_H)]),
    "lzyb_input_open": (C.c_int, [C.c_char_p, C.c_int, C.c_uint64, C.POINTER(_H)]),
    "lzyb_transfer_chunk": (C.c_int64, [_H, _H, C.c_uint64, C.c_uint64, _P]),
    "lzyb_input_finish": (C.c_int, [_H, _H, _P]),
    "lzyb_input_write_host": (C.c_int64, [_H, C.c_uint64, _P, C.c_uint64, _P]),
    "lzyb_input_finish_expect": (C.c_int, [_H, C.POINTER(C.c_uint64 * 2), _P]),
    "lzyb_input_adopt_external": (C.c_int, [_H, C.POINTER(C.c_uint64 * 2), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), _P]),
    "lzyb_input_to_output": (C.c_int, [_H]),
    "lzyb_input_reset": (C.c_int, [_H]),
    "lzyb_slot_export": (C.c_int, [_H, C.POINTER(IpcDesc)]),
    "lzyb_peer_import": (C.c_int, [C.POINTER(IpcDesc), C.c_int, C.POINTER(_H)]),
    "lzyb_sink_export": (C.c_int, [_H, C.POINTER(IpcDesc)]),
    "lzyb_slot_push": (C.c_int, [_H, C.POINTER(IpcDesc), C.c_uint64, C.c_uint64, _P]),
    "lzyb_input_expect_push": (C.c_int64, [_H, C.POINTER(IpcDesc), C.c_uint64, C.c_uint64, _P]),
    "lzyb_content_key": (C.c_int, [C.POINTER(C.c_uint64 * 2), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64 * 2)]),
    "lzyb_broadcast": (C.c_int, [_H, C.POINTER(_H), C.c_int, C.c_uint64]),
    "lzyb_frame_host_object": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "lzyb_ring_read": (C.c_int64, [C.c_uint64, _P, C.c_uint64]),
    "lzyb_ring_frame_info": (C.c_int, [C.c_uint64, C.POINTER(FrameHeader)]),
    "lzyb_ring_release": (C.c_int, [C.c_uint64]),
    "lzyb_digest_device": (C.c_int, [C.c_int, _P, C.c_uint64, C.POINTER(C.c_uint64 * 2), _P]),
    "lzyb_fill_pattern": (C.c_int, [C.c_int, _P, C.c_uint64, C.c_uint64, _P]),
    "lzyb_copy_to_host": (C.c_int, [_H, C.c_uint64, _P, C.c_uint64, _P]),
    "lzyb_stats": (C.c_int, [C.POINTER(Stats)]),
    "lzyb_stats_reset": (C.c_int, []),
    "lzyb_set_copy_engine": (C.c_int, [C.c_int]),
    "lzyb_set_push_min_bytes": (C.c_int, [C.c_uint64]),
    "lzyb_set_grid": (C.c_int, [C.c_int]),
    "lzyb_set_tma_config": (C.c_int, [C.c_int]),
    "lzyb_set_dma_schedule": (C.c_int, [C.c_uint64, C.c_int]),
    "lzyb_dma_schedule": (C.c_int, [C.c_uint64, C.POINTER(C.c_uint64), C.c_int]),
    "lzyb_set_debug_flags": (C.c_int, [C.c_int]),
    "lzyb_fault_inject": (C.c_int, [_H, C.c_int, C.c_uint64]),
}
for _name, (_res, _args) in _SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header and library disagree
    _fn.restype = _res
    _fn.argtypes = _args

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def last_error() -> str:
    return (lib.lzyb_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> int:
    """Raise for negative statuses other than EOF; return rc otherwise."""
    if rc >= 0 or rc == EOF:
        return rc
    msg = last_error() or (lib.lzyb_strerror(rc) or b"").decode()
    if rc == E_PEER:
        raise ReadException(rc, msg)
    if rc == E_NOGPU:
        raise NoGpuError(rc, msg)
    raise LzybError(rc, msg)


def init(first_gpu: int = 0, n_gpus: int = 0, host_ring_bytes: int = 0) -> int:
    check(lib.lzyb_init(first_gpu, n_gpus, host_ring_bytes))
    return lib.lzyb_device_count()


def is_initialized() -> bool:
    return bool(lib.lzyb_is_initialized())


def shutdown() -> None:
    check(lib.lzyb_shutdown())


def stats() -> dict:
    s = Stats()
    check(lib.lzyb_stats(C.byref(s)))
    return s.as_dict()


def digest_hex(d: Sequence[int]) -> str:
    """32 hex characters, the same width as the reference's md5 data hash
    (pylzy/lzy/api/v1/snapshot.py:154-155)."""
    return f"{int(d[0]):016x}{int(d[1]):016x}"


class SlotHandle:
    """Thin object wrapper over one lzyb handle (an HBM frame)."""

    __slots__ = ("h", "slot_id", "gpu", "nbytes", "_closed", "_keepalive", "__weakref__")

    def __init__(self, h: int, slot_id: str, gpu: int, nbytes: int):
        self.h, self.slot_id, self.gpu, self.nbytes = h, slot_id, gpu, nbytes
        self._closed = False
        self._keepalive = None

    # -- construction -----------------------------------------------------------------------
    @classmethod
    def create_output(cls, slot_id: str, gpu: int, nbytes: int) -> "SlotHandle":
        h = _H()
        check(lib.lzyb_slot_create(slot_id.encode(), gpu, nbytes, C.byref(h)))
        return cls(h.value, slot_id, gpu, nbytes)

    @classmethod
    def open_input(cls, slot_id: str, gpu: int, nbytes: int) -> "SlotHandle":
        h = _H()
        check(lib.lzyb_input_open(slot_id.encode(), gpu, nbytes, C.byref(h)))
        return cls(h.value, slot_id, gpu, nbytes)

    @classmethod
    def import_peer(cls, desc: IpcDesc, onto_gpu: int) -> "SlotHandle":
        h = _H()
        check(lib.lzyb_peer_import(C.byref(desc), onto_gpu, C.byref(h)))
        return cls(h.value, f"ipc:{desc.pid}:{desc.gpu}", onto_gpu, int(desc.nbytes))

    # -- output side ------------------------------------------------------------------------
    def device_ptr(self) -> int:
        p = _P()
        check(lib.lzyb_slot_device_ptr(self.h, C.byref(p)))
        return p.value or 0

    def set_meta(self, dtype: int, shape: Sequence[int]) -> None:
        arr = (C.c_uint64 * max(1, len(shape)))(*[int(x) for x in shape])
        check(lib.lzyb_slot_set_meta(self.h, dtype, len(shape), arr))

    def pack(self, src_dptr: int, nbytes: int, stream: int = 0) -> None:
        check(lib.lzyb_slot_pack(self.h, src_dptr, nbytes, stream or None))

    def pack_2d(self, src_dptr: int, rows: int, row_bytes: int, pitch: int, stream: int = 0) -> None:
        check(lib.lzyb_slot_pack_2d(self.h, src_dptr, rows, row_bytes, pitch, stream or None))

    def pack_host(self, src_host_ptr: int, nbytes: int, stream: int = 0) -> None:
        check(lib.lzyb_slot_pack_host(self.h, src_host_ptr, nbytes, stream or None))

    def pack_host_part(self, offset: int, src_host_ptr: int, nbytes: int, last: bool, stream: int = 0) -> None:
        check(lib.lzyb_slot_pack_host_part(self.h, offset, src_host_ptr, nbytes, int(last), stream or None))

    def commit(self, stream: int = 0) -> Tuple[int, int]:
        d = (C.c_uint64 * 2)()
        check(lib.lzyb_slot_commit(self.h, C.byref(d), stream or None))
        return int(d[0]), int(d[1])

    def commit_async(self, stream: int = 0) -> None:
        """commit in stream order (needs ``set_lazy_commit(True)``, else it blocks like ``commit``): the slot is READY for every
        reader that is ordered behind the pack on the device; the digest is harvested by whoever asks first (``commit`` /
        ``info`` / ``export`` / a copying transfer) — include/lzy_b200.h: lzyb_set_lazy_commit"""
        check(lib.lzyb_slot_commit(self.h, None, stream or None))

    def read(self, offset: int, length: int) -> Optional[bytes]:
        """``readFromOffset`` for legacy consumers; None at EOF."""
        buf = C.create_string_buffer(length)
        n = check(lib.lzyb_slot_read(self.h, offset, buf, length))
        return None if n == EOF else buf.raw[:n]

    def copy_to_host(self, dst_ptr: int, length: int, offset: int = 0, stream: int = 0) -> None:
        check(lib.lzyb_copy_to_host(self.h, offset, dst_ptr, length, stream or None))

    # -- input side -------------------------------------------------------------------------
    def transfer_chunk(self, src: "SlotHandle", offset: int, max_len: int, stream: int = 0) -> int:
        """``InputTransfer.transferChunkTo``: bytes moved, or EOF (-1); raises ReadException on
        a peer fault and LzybError on a local one."""
        return check(lib.lzyb_transfer_chunk(self.h, src.h, offset, max_len, stream or None))

    def finish(self, src: "SlotHandle", stream: int = 0) -> None:
        check(lib.lzyb_input_finish(self.h, src.h, stream or None))

    def finish_expect(self, digest: Sequence[int], stream: int = 0) -> None:
        d = (C.c_uint64 * 2)(int(digest[0]), int(digest[1]))
        check(lib.lzyb_input_finish_expect(self.h, C.byref(d), stream or None))

    def write_host(self, offset: int, src_host_ptr: int, length: int, stream: int = 0) -> int:
        return check(lib.lzyb_input_write_host(self.h, offset, src_host_ptr, length, stream or None))

    def adopt_external(self, digest: Optional[Sequence[int]], dtype: int = 0, shape: Sequence[int] = (), stream: int = 0) -> None:
        """the payload was filled in place by an external writer (an NCCL broadcast): scan + verify"""
        d = None if digest is None else C.byref((C.c_uint64 * 2)(int(digest[0]), int(digest[1])))
        arr = (C.c_uint64 * max(1, len(shape)))(*[int(x) for x in shape])
        check(lib.lzyb_input_adopt_external(self.h, d, dtype, len(shape), arr, stream or None))

    def to_output(self) -> None:
        check(lib.lzyb_input_to_output(self.h))

    # -- producer-issued push between processes (include/lzy_b200.h: lzyb_sink_export / lzyb_slot_push / lzyb_input_expect_push)
    def sink_export(self) -> IpcDesc:
        """consumer side: describe this (empty) input frame so that the producer's process can push into it"""
        d = IpcDesc()
        check(lib.lzyb_sink_export(self.h, C.byref(d)))
        return d

    def push(self, sink: IpcDesc, offset: int = 0, length: int = 0, stream: int = 0) -> None:
        """producer side: push_kernel on this slot's GPU, stores landing in the sink frame over NVLink"""
        check(lib.lzyb_slot_push(self.h, C.byref(sink), offset, length, stream or None))

    def expect_push(self, producer: IpcDesc, offset: int = 0, length: int = 0, stream: int = 0) -> int:
        """consumer side: landed_scan_kernel digests the chunks as the producer's push lands them; finish with
        ``finish_expect(producer.digest)``"""
        return check(lib.lzyb_input_expect_push(self.h, C.byref(producer), offset, length, stream or None))

    def set_movable(self, movable: bool = True) -> None:
        """declare a sole consumer: a whole-payload same-GPU transfer from this slot hands the frame over"""
        check(lib.lzyb_slot_set_movable(self.h, int(movable)))

    def reset(self) -> None:
        check(lib.lzyb_input_reset(self.h))

    # -- common -----------------------------------------------------------------------------
    def info(self) -> SlotInfo:
        i = SlotInfo()
        check(lib.lzyb_slot_info(self.h, C.byref(i)))
        return i

    def meta(self) -> SlotInfo:
        """sizes, state, dtype, shape — never waits; the digest fields are zero while the digest is still being computed"""
        i = SlotInfo()
        check(lib.lzyb_slot_meta(self.h, C.byref(i)))
        return i

    def wait(self, stream: int = 0) -> None:
        """order `stream` (0: this thread) after the last launch that wrote the frame"""
        check(lib.lzyb_slot_wait(self.h, stream or None))

    def header(self) -> FrameHeader:
        hd = FrameHeader()
        check(lib.lzyb_slot_header(self.h, C.byref(hd)))
        return hd

    def export(self) -> IpcDesc:
        d = IpcDesc()
        check(lib.lzyb_slot_export(self.h, C.byref(d)))
        return d

    def fault(self, kind: int, arg: int = 0) -> None:
        check(lib.lzyb_fault_inject(self.h, kind, arg))

    def retain(self) -> None:
        check(lib.lzyb_slot_retain(self.h))

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            check(lib.lzyb_slot_close(self.h))

    @classmethod
    def adopt(cls, h: int, owned: bool = True) -> "SlotHandle":
        """wrap a handle produced by the slots runtime (retained when `owned`, so closing is safe)"""
        i = SlotInfo()
        check(lib.lzyb_slot_meta(h, C.byref(i)))
        if owned:
            check(lib.lzyb_slot_retain(h))
        s = cls(h, f"adopted:{h}", int(i.gpu), int(i.nbytes))
        s._closed = not owned
        return s

    def __del__(self):
        try:
            if not self._closed:
                self._closed = True
                lib.lzyb_slot_close(self.h)
        except Exception:
            pass

    def __repr__(self):
        return f"SlotHandle({self.slot_id!r}, gpu={self.gpu}, nbytes={self.nbytes}, h={self.h})"


def broadcast(src: SlotHandle, dsts: Sequence[SlotHandle], chunk_bytes: int = 0) -> None:
    arr = (_H * len(dsts))(*[d.h for d in dsts])
    check(lib.lzyb_broadcast(src.h, arr, len(dsts), chunk_bytes))


def frame_host_object(data: bytes) -> int:
    off = C.c_uint64()
    check(lib.lzyb_frame_host_object(data, len(data), C.byref(off)))
    return off.value


def ring_frame_info(off: int) -> FrameHeader:
    hd = FrameHeader()
    check(lib.lzyb_ring_frame_info(off, C.byref(hd)))
    return hd


def ring_read(off: int) -> bytes:
    hd = ring_frame_info(off)
    buf = C.create_string_buffer(max(1, int(hd.nbytes)))
    n = check(lib.lzyb_ring_read(off, buf, int(hd.nbytes)))
    return buf.raw[:n]


def ring_release(off: int) -> None:
    check(lib.lzyb_ring_release(off))


def digest_device(gpu: int, dptr: int, nbytes: int, stream: int = 0) -> Tuple[int, int]:
    d = (C.c_uint64 * 2)()
    check(lib.lzyb_digest_device(gpu, dptr, nbytes, C.byref(d), stream or None))
    return int(d[0]), int(d[1])


def fill_pattern(gpu: int, dptr: int, nbytes: int, seed: int, stream: int = 0) -> None:
    check(lib.lzyb_fill_pattern(gpu, dptr, nbytes, seed, stream or None))


def set_copy_engine(engine: int) -> None:
    check(lib.lzyb_set_copy_engine(engine))


def set_transfer_waves(waves: int) -> None:
    check(lib.lzyb_set_transfer_waves(int(waves)))


def set_lazy_commit(on: bool) -> None:
    check(lib.lzyb_set_lazy_commit(int(bool(on))))


HOST_BLOCK_MIN, HOST_BLOCK_MAX = 64 << 10, 16 << 20


def pinned_buffer(nbytes: int):
    """A writable ctypes byte array over a pooled pinned host block (lzyb_host_alloc), returned to the pool when the last
    reference to it — or to anything built on it, e.g. a numpy view — goes away; None when the pool declines (size outside
    its classes, limit reached): the caller then uses ordinary memory."""
    if not HOST_BLOCK_MIN <= nbytes <= HOST_BLOCK_MAX:
        return None
    p = C.c_void_p()
    if lib.lzyb_host_alloc(nbytes, C.byref(p)) != OK:
        return None
    buf = (C.c_ubyte * nbytes).from_address(p.value)
    weakref.finalize(buf, lib.lzyb_host_free, p.value)
    return buf


def set_host_pool_limit(nbytes: int) -> None:
    check(lib.lzyb_set_host_pool_limit(nbytes))


def set_push_min_bytes(nbytes: int) -> None:
    check(lib.lzyb_set_push_min_bytes(nbytes))


def content_key(digest: Sequence[int], dtype_flags: int, shape: Sequence[int]) -> Tuple[int, int]:
    """content address of a framed value: the payload digest with (dtype|flags, ndim, shape) folded in
    (include/lzyb_digest.h: lzyb_content_key_fold).  Host arithmetic; works without a GPU."""
    d = (C.c_uint64 * 2)(int(digest[0]), int(digest[1]))
    out = (C.c_uint64 * 2)()
    arr = (C.c_uint64 * max(1, len(shape)))(*[int(x) for x in shape])
    check(lib.lzyb_content_key(C.byref(d), dtype_flags & 0xFFFFFFFF, len(shape), arr, C.byref(out)))
    return int(out[0]), int(out[1])


def set_dma_schedule(last_bytes: int = 0, growth: int = 6) -> None:
    check(lib.lzyb_set_dma_schedule(last_bytes, growth))


def dma_schedule(nbytes: int) -> list:
    buf = (C.c_uint64 * 64)()
    n = lib.lzyb_dma_schedule(nbytes, buf, 64)
    return [int(buf[i]) for i in range(min(n, 64))]


def set_tma_config(index: int) -> None:
    check(lib.lzyb_set_tma_config(index))


def set_grid(ctas_per_sm: int) -> None:
    check(lib.lzyb_set_grid(ctas_per_sm))


def fault_global(kind: int, arg: int = 0) -> None:
    check(lib.lzyb_fault_inject(0, kind, arg))
