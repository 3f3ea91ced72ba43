"""oracle/ — TEST INFRASTRUCTURE ONLY (CPU restatement of the reference's slot/channel data path).

Nothing under ``lzy_b200/`` may import this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py`` (``cpu_baseline`` / ``--impl reference`` legs), and
there only as the checker / the timed CPU baseline, never as the thing shipped.

Parity status:
 * byte-movement semantics (chunked stream, offset resume, FS copy loop, MD5 URI rule) are pinned
   against the reference's own pure-Python leaf modules executed in the authoring container
   (``tests/golden/make_golden.py`` → ``tests/golden/*.json``) and against the reference's
   known-answer tests restated in ``tests/``;
 * the object↔bytes encoding lives in the un-vendored third-party ``serialzy`` (>=1.2,<2,
   pylzy/requirements.txt:19) — its byte format is **parity unpinned**; the oracle uses cloudpickle
   and parity is stated at the value level (dtype/shape/bytes of arrays, ``==`` of objects).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "_build" / "liblzyb_oracle.so"


def build() -> Path:
    subprocess.run(["make", "-C", str(_DIR), "-s"], check=True)
    return _SO


def _load():
    if not _SO.exists():
        build()
    lib = C.CDLL(str(_SO))
    u8p, u64 = C.c_void_p, C.c_uint64
    lib.oracle_stream_copy.restype = u64
    lib.oracle_stream_copy.argtypes = [u8p, u8p, u64, u64, u64, C.POINTER(u64)]
    lib.oracle_fs_copy.restype = u64
    lib.oracle_fs_copy.argtypes = [u8p, u8p, u64]
    lib.oracle_digest_partial.restype = None
    lib.oracle_digest_partial.argtypes = [u8p, u64, u64, C.POINTER(u64), C.POINTER(u64)]
    lib.oracle_digest.restype = None
    lib.oracle_digest.argtypes = [u8p, u64, C.POINTER(u64 * 2)]
    lib.oracle_digest_finalize.restype = None
    lib.oracle_digest_finalize.argtypes = [u64, u64, u64, C.POINTER(u64 * 2)]
    lib.oracle_pattern.restype = None
    lib.oracle_pattern.argtypes = [u8p, u64, u64]
    lib.oracle_gather2d.restype = None
    lib.oracle_gather2d.argtypes = [u8p, u8p, u64, u64, u64]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _ptr(a: np.ndarray) -> int:
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def as_bytes_array(x) -> np.ndarray:
    if isinstance(x, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(x), dtype=np.uint8)
    a = np.ascontiguousarray(x)
    return a.view(np.uint8).reshape(-1)


def digest(x) -> tuple:
    a = as_bytes_array(x)
    out = (C.c_uint64 * 2)()
    lib().oracle_digest(_ptr(a) if a.size else None, a.size, C.byref(out))
    return int(out[0]), int(out[1])


def digest_partial(x, byte_offset: int, acc=(0, 0)) -> tuple:
    a = as_bytes_array(x)
    s, xr = C.c_uint64(acc[0]), C.c_uint64(acc[1])
    lib().oracle_digest_partial(_ptr(a) if a.size else None, a.size, byte_offset, C.byref(s), C.byref(xr))
    return int(s.value), int(xr.value)


def digest_finalize(sum_: int, xr: int, nbytes: int) -> tuple:
    out = (C.c_uint64 * 2)()
    lib().oracle_digest_finalize(sum_, xr, nbytes, C.byref(out))
    return int(out[0]), int(out[1])


def digest_hex(x) -> str:
    d = digest(x)
    return f"{d[0]:016x}{d[1]:016x}"


def pattern(nbytes: int, seed: int) -> np.ndarray:
    a = np.empty(nbytes, dtype=np.uint8)
    lib().oracle_pattern(_ptr(a) if nbytes else None, nbytes, seed)
    return a


def stream_copy(src: np.ndarray, dst: np.ndarray, offset: int = 0, fail_after: int = -1) -> tuple:
    """(chunks sent, bytes delivered): OutputSlot.ReadThread → SlotInputTransfer.transferChunkTo."""
    delivered = C.c_uint64()
    fa = (1 << 64) - 1 if fail_after < 0 else fail_after
    n = lib().oracle_stream_copy(_ptr(src), _ptr(dst), src.size, offset, fa, C.byref(delivered))
    return int(n), int(delivered.value)


def fs_copy(src: np.ndarray, dst: np.ndarray) -> int:
    return int(lib().oracle_fs_copy(_ptr(src), _ptr(dst), src.size))


def gather2d(src: np.ndarray, rows: int, row_bytes: int, pitch: int) -> np.ndarray:
    dst = np.empty(rows * row_bytes, dtype=np.uint8)
    lib().oracle_gather2d(_ptr(src), _ptr(dst), rows, row_bytes, pitch)
    return dst
