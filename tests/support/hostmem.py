"""Test-only: builds and installs the in-memory slot backend (tests/support/hostmem_backend.c) at the
backend seam of the C++ slots runtime — the analogue of SlotsTest.InMemBackend."""
import ctypes as C
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
_SO = Path(__file__).resolve().parent / "_build" / "libhostmem_backend.so"
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = Path(__file__).resolve().parent / "hostmem_backend.c"
    libdir = ROOT / "lzy_b200" / "lib"
    deps = [src, ROOT / "include" / "lzy_b200.h", ROOT / "include" / "lzy_b200_host.h"]  # the ops table is ABI
    if not _SO.exists() or _SO.stat().st_mtime < max(d.stat().st_mtime for d in deps):
        _SO.parent.mkdir(exist_ok=True)
        subprocess.run(["gcc", "-O1", "-fPIC", "-shared", "-Wall", "-o", str(_SO), str(src), f"-L{libdir}",
                        "-llzy_b200", f"-Wl,-rpath,{libdir}", "-lpthread", "-lrt"], check=True)
    from lzy_b200 import _native  # noqa: F401  (liblzy_b200.so must be loaded first)
    lib = C.CDLL(str(_SO))
    lib.hostmem_write.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.hostmem_peek.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.hostmem_peek.restype = C.c_int64
    lib.hostmem_moving_calls.restype = C.c_longlong
    lib.hostmem_size.argtypes = [C.c_uint64]
    lib.hostmem_size.restype = C.c_int64
    lib.hostmem_set_fail_after.argtypes = [C.c_uint64, C.c_longlong]
    lib.hostmem_set_fail_read.argtypes = [C.c_uint64, C.c_int]
    lib.hostmem_corrupt.argtypes = [C.c_uint64, C.c_uint64]
    _lib = lib
    return lib


def install():
    from lzy_b200 import _host
    lib = load()
    ops = C.c_void_p.in_dll(lib, "hostmem_backend_ops")
    _host.lib.lzyb_slots_set_backend(C.addressof(ops))
    return lib


def uninstall():
    from lzy_b200 import _host
    _host.lib.lzyb_slots_set_backend(None)


def write(h, data: bytes, offset=0):
    assert load().hostmem_write(h, offset, data, len(data)) == 0


def peek(h, n, offset=0) -> bytes:
    buf = C.create_string_buffer(max(1, n))
    got = load().hostmem_peek(h, offset, buf, n)
    assert got >= 0
    return buf.raw[:got]


def info(h):
    from lzy_b200 import _native as N
    out = N.SlotInfo()
    assert load().hostmem_info(C.c_uint64(h), C.byref(out)) == 0
    return out


def set_meta(h, dtype: int, shape) -> None:
    arr = (C.c_uint64 * max(1, len(shape)))(*shape)
    assert load().hostmem_set_meta(C.c_uint64(h), C.c_uint32(dtype), C.c_uint32(len(shape)), arr) == 0


def size(h) -> int:
    n = load().hostmem_size(h)
    assert n >= 0
    return int(n)


class HostmemMount:
    """Test stand-in for `lzy_b200.api.v1.startup.FrameMount`: slot paths resolve to the in-memory backend's
    slots of one execution context and objects cross as their serializers' byte streams."""

    def __init__(self, ctx):
        self.ctx = ctx

    def read(self, path, typ, serializers):
        from io import BytesIO
        h = self.ctx.input_handle(path)
        ser = serializers.find_serializer_by_type(typ)
        return ser.deserialize(BytesIO(peek(h, size(h))), typ if isinstance(typ, type) else None)

    def write(self, path, typ, data, serializers):
        from io import BytesIO
        ser = serializers.find_serializer_by_type(typ)
        buf = BytesIO()
        ser.serialize(data, buf)
        raw = buf.getvalue()
        h = self.ctx.output_open(path, len(raw))
        write(h, raw)
        self.ctx.output_done(path)
