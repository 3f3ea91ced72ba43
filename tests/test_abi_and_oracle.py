"""CPU-side checks: the C-ABI library loads and exports every symbol the headers declare (no compute
without a GPU), the product refuses to run without CUDA, and the oracle's C restatement agrees with
its own definition and with the reference's pure-Python leaf code (golden vectors)."""
import ctypes
import hashlib
import json
import re
from pathlib import Path

import numpy as np
import pytest

import oracle

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols(header: Path):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(lzyb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from lzy_b200 import _host, _native
    lib = ctypes.CDLL(str(ROOT / "lzy_b200" / "lib" / "liblzy_b200.so"))
    for header, bound in (("lzy_b200.h", _native.EXPORTED_SYMBOLS), ("lzy_b200_host.h", _host.EXPORTED_SYMBOLS)):
        syms = declared_symbols(ROOT / "include" / header)
        assert len(syms) > 20
        for s in syms:
            assert hasattr(lib, s), f"{s} declared in {header} but not exported"
        assert sorted(bound) == syms, f"python binding and {header} disagree: {set(bound) ^ set(syms)}"


def test_no_cpu_fallback_without_gpu():
    import torch
    from lzy_b200 import _native as N
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    with pytest.raises(N.NoGpuError):
        N.init(0, 1, 0)
    h = ctypes.c_uint64()
    rc = N.lib.lzyb_slot_create(b"x", 0, 16, ctypes.byref(h))
    assert rc == N.E_STATE and "lzyb_init" in N.last_error()
    off = ctypes.c_uint64()
    assert N.lib.lzyb_frame_host_object(b"abc", 3, ctypes.byref(off)) == N.E_STATE


def test_product_never_imports_oracle():
    for py in (ROOT / "lzy_b200").rglob("*.py"):
        src = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{py} imports the oracle"
    for src in (ROOT / "lzy_b200" / "csrc").rglob("*"):
        if src.is_file() and src.suffix in (".cu", ".cuh", ".cpp", ".h"):
            assert "oracle/" not in src.read_text(), f"{src} references the oracle"


def py_digest(data: bytes):
    """independent pure-Python statement of include/lzyb_digest.h (small inputs only)"""
    M = (1 << 64) - 1
    K_POS, K1, K2, K_LEN = 0x9E3779B97F4A7C15, 0xFF51AFD7ED558CCD, 0xC4CEB9FE1A85EC53, 0xD6E8FEB86659FD93

    def mix(x):
        x = (x * K1) & M
        x ^= x >> 32
        x = (x * K2) & M
        x ^= x >> 29
        return x

    s = xr = 0
    for i in range(0, len(data), 8):
        w = int.from_bytes(data[i:i + 8].ljust(8, b"\0"), "little")
        m = mix(w ^ (((i // 8 + 1) * K_POS) & M))
        s = (s + m) & M
        xr ^= m
    ln = ((len(data) + 1) * K_LEN) & M
    hi = mix(s ^ ln)
    lo = mix((xr + (((s << 32) | (s >> 32)) & M) + ln) & M)
    return hi, lo ^ (hi >> 1)


@pytest.mark.parametrize("n", [0, 1, 5, 8, 9, 16, 17, 1000, 4099])
def test_c_oracle_digest_matches_pure_python_definition(n):
    data = bytes(np.random.default_rng(n).integers(0, 256, n, dtype=np.uint8))
    assert oracle.digest(data) == py_digest(data)


def test_digest_golden_vectors():
    g = json.loads((ROOT / "tests" / "golden" / "digest_vectors.json").read_text())
    for v in g["vectors"]:
        data = bytes.fromhex(v["hex"]) if "hex" in v else oracle.pattern(v["pattern_nbytes"], v["pattern_seed"]).tobytes()
        assert oracle.digest_hex(data) == v["digest"], v.get("name")


def test_digest_detects_small_integer_permutations():
    """the collision class that rules out a linear (Fletcher-style) checksum for content addressing"""
    a = np.array([0, 2, 0], dtype=np.int32)
    b = np.array([1, 0, 1], dtype=np.int32)
    assert oracle.digest(a) != oracle.digest(b)
    x = np.arange(64, dtype=np.uint64)
    y = x.copy()
    y[[3, 40]] = y[[40, 3]]
    assert oracle.digest(x) != oracle.digest(y)


def py_content_key(digest, dtype_flags, shape):
    """independent pure-Python statement of lzyb_content_key_fold (include/lzyb_digest.h)"""
    M = (1 << 64) - 1
    K_POS, K1, K2, K_LEN, K_META = (0x9E3779B97F4A7C15, 0xFF51AFD7ED558CCD, 0xC4CEB9FE1A85EC53, 0xD6E8FEB86659FD93,
                                    0xA0761D6478BD642F)

    def mix(x):
        x = (x * K1) & M
        x ^= x >> 32
        x = (x * K2) & M
        x ^= x >> 29
        return x

    h = mix((((dtype_flags & 0xFFFFFFFF) << 32) | len(shape)) ^ K_META)
    for i, d in enumerate(shape):
        h = mix(h ^ mix((d + (i + 1) * K_POS) & M))
    hi = mix(digest[0] ^ h)
    lo = mix((digest[1] + (((h << 32) | (h >> 32)) & M) + K_LEN) & M)
    return hi, lo ^ (hi >> 1)


def test_content_key_tells_apart_values_that_share_their_bytes():
    """the payload digest is blind to dtype / shape / serializer flags, so zeros(4, f32) and
    zeros(4, i32) or one buffer seen as (2, 8) and (4, 4) had one URI.  The reference's md5 runs over the serialized
    stream, which carries both (snapshot.py:147-155).  The content key folds the header metadata in."""
    from lzy_b200 import _native as N
    d = oracle.digest(np.zeros(16, dtype=np.uint8))
    cases = [(N.DT_F32, (4,)), (N.DT_I32, (4,)), (N.DT_F32, (2, 2)), (N.DT_F32, (1, 4)), (N.DT_F32, (4, 1)),
             (N.DT_F32 | (1 << 16), (4,)), (N.DT_PICKLE, (16,)), (N.DT_U8, (16,)), (N.DT_U8, (2, 8)), (N.DT_U8, (8, 2)),
             (N.DT_U8, (4, 4)), (N.DT_U8, ())]
    keys = [N.content_key(d, dt, sh) for dt, sh in cases]
    assert len(set(keys)) == len(keys)
    for (dt, sh), k in zip(cases, keys):
        assert k == py_content_key(d, dt, sh)
        assert k == N.content_key(d, dt, sh)  # a function of its inputs only
    # and it still depends on every bit of the payload digest
    assert N.content_key((d[0] ^ 1, d[1]), N.DT_F32, (4,)) != keys[0]
    assert N.content_key((d[0], d[1] ^ (1 << 63)), N.DT_F32, (4,)) != keys[0]


def test_partial_digests_compose():
    a = np.random.default_rng(3).integers(0, 256, 5 * (1 << 20) + 11, dtype=np.uint8)
    acc = (0, 0)
    for off in range(0, a.size, 2 << 20):
        acc = oracle.digest_partial(a[off:off + (2 << 20)], off, acc)
    assert oracle.digest_finalize(acc[0], acc[1], a.size) == oracle.digest(a)


def test_stream_copy_oracle_chunking_and_offset_resume():
    """OutputSlot.ReadThread: ≤ 2 MiB chunks from `offset`; SlotInputTransfer writes at the running
    offset; testRestoreReadFromOffset: fail after 1 MiB, resume, equal."""
    n = 5 * (1 << 20) + 123
    src = np.random.default_rng(0).integers(0, 256, n, dtype=np.uint8)
    dst = np.zeros(n, dtype=np.uint8)
    chunks, delivered = oracle.stream_copy(src, dst)
    assert (chunks, delivered) == (3, n) and np.array_equal(src, dst)
    dst[:] = 0
    chunks, delivered = oracle.stream_copy(src, dst, 0, fail_after=1 << 20)
    assert delivered == 2 << 20  # one 2 MiB read() had already been issued when the stream noticed
    assert np.array_equal(dst[:delivered], src[:delivered]) and not dst[delivered:].any()
    chunks, delivered = oracle.stream_copy(src, dst, delivered)
    assert delivered == n and np.array_equal(src, dst)


def test_reference_leaf_module_golden_vectors():
    """Outputs recorded from the reference's own pure-Python modules (hashing.py, fs.py), see
    tests/golden/make_golden.py; the oracle's restatements must reproduce them."""
    from oracle import local_runtime as L
    g = json.loads((ROOT / "tests" / "golden" / "reference_leaf_vectors.json").read_text())
    for v in g["md5_of_str"]:
        assert L.md5_of_str(v["in"]) == v["out"]
    for v in g["hashing_io"]:
        import io
        sink = io.BytesIO()
        h = L.HashingIO(sink)
        for part in v["writes_hex"]:
            h.write(bytes.fromhex(part))
        assert h.md5 == v["md5"] and sink.getvalue().hex() == "".join(v["writes_hex"])
    for v in g["fs_copy"]:
        data = oracle.pattern(v["nbytes"], v["seed"]).tobytes()
        assert hashlib.md5(data).hexdigest() == v["src_md5"]
        import tempfile, os, io
        with tempfile.TemporaryDirectory() as d:
            uri = "file://" + os.path.join(d, "a", "b", "blob")
            L.fs_write(uri, io.BytesIO(data))
            out = io.BytesIO()
            L.fs_read(uri, out)
            assert hashlib.md5(out.getvalue()).hexdigest() == v["dst_md5"] == v["src_md5"]
            assert L.path_from_uri(uri) == os.path.join(d, "a", "b", "blob")
        assert oracle.fs_copy(np.frombuffer(data, np.uint8), np.empty(len(data), np.uint8)) == v["loop_iterations"]
    for v in g["path_from_uri"]:
        assert L.path_from_uri(v["in"]) == v["out"]


def test_copy_engine_pull_schedule_arithmetic():
    """launch_dma_verify's sub-chunk plan (capi.cu: dma_schedule) is pure arithmetic and callable without a GPU:
    the pieces tile the chunk exactly, every boundary is 16-byte aligned (digest word positions), the ragged tail
    rides in the last piece, pieces shrink geometrically towards the end, and there are only a few of them."""
    import random

    from lzy_b200 import _native as N
    try:
        assert N.dma_schedule(1 << 30) == [800 << 20, 192 << 20, 32 << 20]
        assert N.dma_schedule(4 << 30) == [2720 << 20, 1152 << 20, 192 << 20, 32 << 20]
        assert N.dma_schedule(256 << 20) == [200 << 20, 48 << 20, 8 << 20]
        assert N.dma_schedule(100) == [100] and N.dma_schedule(0) == [0]
        rnd = random.Random(5)
        for _ in range(300):
            n = rnd.choice([rnd.randrange(1, 1 << 12), rnd.randrange(1, 1 << 26), rnd.randrange(1 << 26, 1 << 34)])
            sizes = N.dma_schedule(n)
            assert sum(sizes) == n and 1 <= len(sizes) <= 8
            assert all(s % 16 == 0 for s in sizes[:-1]) and sizes[-1] % 16 == n % 16
            assert all(a >= b for a, b in zip(sizes[1:-1], sizes[2:]))  # geometric tail; the first piece takes the rest
        N.check(N.lib.lzyb_set_dma_schedule(3 << 20, 1))  # uniform 3 MiB pieces (tests use this to force many copies)
        assert N.dma_schedule(8 << 20) == [2 << 20, 3 << 20, 3 << 20]
        assert N.dma_schedule((40 << 20) + 11)[-1] == (3 << 20) + 11
        with pytest.raises(N.LzybError):
            N.check(N.lib.lzyb_set_dma_schedule(100, 6))
    finally:
        N.lib.lzyb_set_dma_schedule(0, 6)


def test_oracle_snapshot_matches_the_reference_snapshot_on_a_real_directory(tmp_path):
    """tests/golden/reference_snapshot_vectors.json: pylzy's own DefaultSnapshot over its own FsStorageClient (executed
    in the authoring container) — the oracle's restatement, which the CPU baseline and the local-path parity tests
    run on, gives the same content addresses (md5 for md5), stores equal data once and reads back the same"""
    import importlib.util
    import json
    import os
    from pathlib import Path

    from oracle import local_runtime

    golden = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("snapshot_battery", golden / "snapshot_battery.py")
    battery = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(battery)
    prefix, inputs = f"file://{tmp_path}/lzy_runs/wf/inputs/", f"{tmp_path}/lzy_runs/wf/inputs"
    snap = local_runtime.OracleSnapshot(prefix)

    def get(entry):
        found, value = snap.get_data(entry.id)
        return ("result", repr(value)) if found else ("absence", "")

    ours = battery.run(dict(
        create=snap.create_entry, put=lambda entry, value: snap.put_data(entry.id, value), get=get,
        uri=lambda entry: entry.storage_uri, hash=lambda entry: entry.data_hash,
        set_uri=lambda entry, uri: setattr(entry, "storage_uri", uri),
        stored=lambda: sorted(os.listdir(inputs)) if os.path.isdir(inputs) else []), prefix)
    assert json.loads(json.dumps(ours)) == json.loads((golden / "reference_snapshot_vectors.json").read_text())


def test_oracle_runs_calls_in_the_order_the_reference_local_runtime_does():
    """tests/golden/reference_order_vectors.json: execution orders of pylzy's own LocalRuntime.exec (run in the
    authoring container with the per-call execution replaced by a recorder) for forty seeded random graphs"""
    import importlib.util
    import json
    from pathlib import Path

    from oracle.local_runtime import OracleLocalRuntime

    golden = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("order_battery", golden / "order_battery.py")
    battery = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(battery)
    ours = battery.run(dict(order=lambda calls: [c.id for c in OracleLocalRuntime.topo_order(calls)]))
    assert ours == json.loads((golden / "reference_order_vectors.json").read_text())
