"""Property-based checks (hypothesis) of the pieces whose contracts are algebraic: the digest composes over any
8-byte-aligned chunking and any chunk order, differs when a byte or the length differs, and agrees between the C
restatement and the pure-Python definition; a stream with offset resume delivers exactly the source; the request
codec and the File byte format round-trip arbitrary values; the copy-engine sub-chunk plan always tiles its chunk."""
import os
import tempfile
from io import BytesIO

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle
from lzy_b200 import _native as N
from lzy_b200.serialization.registry import LzySerializerRegistry
from lzy_b200.types import File
from lzy_b200.utils.argv_codec import decode, encode

from tests.test_abi_and_oracle import py_digest  # the independent pure-Python statement of include/lzyb_digest.h

payloads = st.binary(min_size=0, max_size=4096)


@settings(max_examples=150, deadline=None)
@given(payloads)
def test_c_digest_equals_the_python_definition(data):
    assert oracle.digest(np.frombuffer(data, dtype=np.uint8)) == py_digest(data)


@settings(max_examples=150, deadline=None)
@given(payloads, st.lists(st.integers(min_value=0, max_value=512), max_size=6), st.randoms(use_true_random=False))
def test_digest_composes_over_any_aligned_chunking_in_any_order(data, cuts, rnd):
    a = np.frombuffer(data, dtype=np.uint8)
    bounds = sorted({0, len(data), *[min(len(data), 8 * c) for c in cuts]})
    chunks = list(zip(bounds[:-1], bounds[1:]))
    rnd.shuffle(chunks)
    acc = (0, 0)
    for lo, hi in chunks:
        acc = oracle.digest_partial(a[lo:hi], lo, acc)
    assert oracle.digest_finalize(acc[0], acc[1], len(data)) == oracle.digest(a)


@settings(max_examples=150, deadline=None)
@given(st.binary(min_size=1, max_size=2048), st.data())
def test_digest_sees_every_byte_its_position_and_the_length(data, draw):
    a = np.frombuffer(data, dtype=np.uint8)
    base = oracle.digest(a)
    i = draw.draw(st.integers(min_value=0, max_value=len(data) - 1))
    flipped = bytearray(data)
    flipped[i] ^= 1 << draw.draw(st.integers(min_value=0, max_value=7))
    assert oracle.digest(np.frombuffer(bytes(flipped), dtype=np.uint8)) != base
    assert oracle.digest(np.frombuffer(data + b"\0", dtype=np.uint8)) != base  # zero padding is not free
    if len(data) >= 16 and data[:8] != data[8:16]:
        swapped = data[8:16] + data[:8] + data[16:]
        assert oracle.digest(np.frombuffer(swapped, dtype=np.uint8)) != base   # order matters


@settings(max_examples=100, deadline=None)
@given(st.binary(min_size=0, max_size=3 << 16), st.integers(min_value=0, max_value=3 << 16))
def test_stream_with_a_failure_and_offset_resume_delivers_the_source(data, fail_at):
    src = np.frombuffer(data, dtype=np.uint8)
    dst = np.zeros(len(data), dtype=np.uint8)
    fail_at = min(fail_at, len(data))
    chunks, position = oracle.stream_copy(src, dst, 0, fail_after=fail_at)  # (2 MiB chunks sent, bytes delivered so far)
    assert position <= len(data) and np.array_equal(dst[:position], src[:position])
    assert chunks == -(-position // (2 << 20))
    _, end = oracle.stream_copy(src, dst, position)  # the consumer reopens at the offset it holds
    assert end == len(data) and np.array_equal(dst, src)


values = st.recursive(st.none() | st.booleans() | st.integers() | st.floats(allow_nan=False) | st.text() | st.binary(),
                      lambda inner: st.lists(inner, max_size=4) | st.tuples(inner, inner) | st.dictionaries(st.text(max_size=5), inner, max_size=4),
                      max_leaves=12)


@settings(max_examples=150, deadline=None)
@given(values)
def test_request_codec_round_trips_and_stays_one_shell_word(value):
    text = encode(value)
    assert text.isascii() and not any(c.isspace() for c in text) and decode(text) == value


def test_request_codec_refuses_what_is_not_base64():
    for bad in ("not base64!", "AAA", "ʎzy"):
        with pytest.raises((ValueError, UnicodeEncodeError)):
            decode(bad)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(st.binary(max_size=20000), st.integers(min_value=0, max_value=0o777))
def test_file_serializer_round_trips_content_and_mode(content, mode):
    ser = LzySerializerRegistry().find_serializer_by_type(File)
    with tempfile.NamedTemporaryFile(delete=False) as f:
        f.write(content)
    try:
        os.chmod(f.name, mode | 0o400)  # the serializer has to be able to read it back
        buf = BytesIO()
        ser.serialize(File(f.name), buf)
        assert buf.getvalue() == b"permissions" + oct(mode | 0o400).encode() + content
        buf.seek(0)
        back = ser.deserialize(buf, File)
        try:
            with back.open("rb") as fh:
                assert fh.read() == content
            assert back.stat().st_mode & 0o777 == mode | 0o400
        finally:
            os.chmod(str(back), 0o600)
            os.unlink(str(back))
    finally:
        os.chmod(f.name, 0o600)
        os.unlink(f.name)


@settings(max_examples=300, deadline=None)
@given(st.integers(min_value=0, max_value=1 << 36))
def test_copy_engine_plan_always_tiles_its_chunk(nbytes):
    sizes = N.dma_schedule(nbytes)
    assert sum(sizes) == nbytes and 1 <= len(sizes) <= 8
    assert all(s % 16 == 0 for s in sizes[:-1])
    assert nbytes < (8 << 20) or sizes[-1] <= (32 << 20) + 15  # a short tail scan


_DTYPES = ["uint8", "int8", "int16", "int32", "int64", "float16", "float32", "float64", "bool"]


@settings(max_examples=80, deadline=None)
@given(st.sampled_from(_DTYPES), st.lists(st.integers(min_value=0, max_value=7), min_size=0, max_size=4), st.integers(0, 2 ** 31))
def test_array_stream_form_round_trips_dtype_shape_and_bytes(dtype, shape, seed):
    """the byte-stream form of the frame serializers (what storage holds when a payload is not in HBM): header +
    C-order bytes, for ndarray and for CPU torch tensors"""
    import torch
    reg = LzySerializerRegistry()
    rng = np.random.default_rng(seed)
    a = (rng.integers(0, 2, shape).astype(bool) if dtype == "bool" else
         rng.integers(-100, 100, shape).astype(dtype) if "int" in dtype else rng.standard_normal(shape).astype(dtype))
    ser = reg.find_serializer_by_type(np.ndarray)
    buf = BytesIO()
    ser.serialize(a, buf)
    buf.seek(0)
    back = ser.deserialize(buf, np.ndarray)
    assert back.dtype == a.dtype and back.shape == a.shape and back.tobytes() == a.tobytes()

    t = torch.from_numpy(np.ascontiguousarray(a))
    tser = reg.find_serializer_by_type(torch.Tensor)
    buf = BytesIO()
    tser.serialize(t, buf)
    buf.seek(0)
    tback = tser.deserialize(buf, torch.Tensor)
    assert tback.dtype == t.dtype and tuple(tback.shape) == tuple(t.shape) and torch.equal(tback, t)


# ---- argument binding agrees with Python's own -----------------------------------------------------------------------
@settings(max_examples=200, deadline=None)
@given(st.integers(0, 3), st.integers(0, 3), st.booleans(), st.booleans(), st.integers(0, 5),
       st.lists(st.sampled_from(["p0", "p1", "p2", "zz"]), unique=True, max_size=3))
def test_call_binding_accepts_exactly_what_python_would(n_params, n_defaults, varargs, varkw, n_args, kw_names):
    """the arity rules of infer_and_validate_call_signature (pylzy/lzy/core/call.py:266-287: KeyError for a missing,
    an unexpected positional or an unknown keyword argument) against inspect.Signature.bind on generated functions;
    a parameter given both ways is left to the op's own call, as in the reference"""
    import inspect
    from types import SimpleNamespace

    from lzy_b200.core.call import infer_and_validate_call_signature
    from lzy_b200.serialization.registry import LzySerializerRegistry

    n_defaults = min(n_defaults, n_params)
    params = [f"p{i}: int" + (" = 0" if i >= n_params - n_defaults else "") for i in range(n_params)]
    if varargs:
        params.append("*rest")
    if varkw:
        params.append("**named")
    scope: dict = {}
    exec(f"def f({', '.join(params)}) -> int:\n    return 0\n", scope)
    f = scope["f"]
    args = tuple(range(n_args))
    kwargs = {name: 1 for name in kw_names}
    if any(name in [f"p{i}" for i in range(min(n_args, n_params))] for name in kwargs):
        return  # given both positionally and by name
    try:
        inspect.signature(f).bind(*args, **kwargs)
        python_accepts = True
    except TypeError:
        python_accepts = False
    snapshot = SimpleNamespace(get=lambda eid: None)
    try:
        sig = infer_and_validate_call_signature(f, *args, output_type=(int,), snapshot=snapshot,
                                                serializer_registry=LzySerializerRegistry(), **kwargs)
        ours_accepts = True
    except KeyError:
        ours_accepts = False
    assert ours_accepts == python_accepts, (params, args, kwargs)
    if ours_accepts:
        assert sig.args == args and sig.kwargs == kwargs
        assert len(sig.func.arg_names) == n_args and list(sig.func.arg_names[:n_params]) == [f"p{i}" for i in range(min(n_args, n_params))]
        assert set(sig.func.input_types) == set(sig.func.arg_names) | set(kwargs)


# ---- the dataflow planner ------------------------------------------------------------------------------------------
@settings(max_examples=200, deadline=None)
@given(st.lists(st.tuples(st.booleans(), st.lists(st.integers(0, 6), unique=True, max_size=3)), max_size=8),
       st.sets(st.integers(0, 6), max_size=3))
def test_dataflow_plan_invariants(call_specs, stored):
    """lzy-service DataFlowGraph.java:51-72 as invariants: one supplier per URI; a dropped duplicate's results all
    point at kept suppliers of the very same URIs; cached calls are exactly the cache=True calls whose results are
    all stored; order of the kept calls is submission order; partial overlaps are refused"""
    from types import SimpleNamespace

    from lzy_b200.core.dataflow import plan_dataflow

    uri_of, calls = {}, []
    for i, (cache, uris) in enumerate(call_specs):
        eids = []
        for j, u in enumerate(uris):
            eid = f"c{i}e{j}"
            uri_of[eid] = f"uri-{u}"
            eids.append(eid)
        calls.append(SimpleNamespace(cache=cache, entry_ids=eids, n=i))
    exists = lambda u: int(u.split("-")[1]) in stored  # noqa: E731

    # what must happen, restated naively
    supplied, expect_error = set(), False
    for c in calls:
        us = [uri_of[e] for e in c.entry_ids]
        if c.cache and us and all(exists(u) for u in us):
            continue
        if us and all(u in supplied for u in us):
            continue
        if any(u in supplied for u in us):
            expect_error = True
            break
        supplied.update(us)
    if expect_error:
        with pytest.raises(RuntimeError, match="already exists"):
            plan_dataflow(calls, uri_of.__getitem__, exists)
        return

    plan = plan_dataflow(calls, uri_of.__getitem__, exists)
    assert sorted(c.n for c in plan.calls + plan.cached + plan.duplicates) == list(range(len(calls)))
    assert [c.n for c in plan.calls] == sorted(c.n for c in plan.calls)
    for c in plan.cached:
        assert c.cache and c.entry_ids and all(exists(uri_of[e]) for e in c.entry_ids)
    kept_uris = [uri_of[e] for c in plan.calls for e in c.entry_ids]
    assert len(kept_uris) == len(set(kept_uris)), "two suppliers for one URI"
    for i, c in enumerate(plan.calls):
        assert all(plan.supplier[e] == i for e in c.entry_ids)
    for d in plan.duplicates:
        for e in d.entry_ids:
            supplier = plan.calls[plan.supplier[e]]
            assert uri_of[e] in {uri_of[x] for x in supplier.entry_ids} and supplier.n < d.n
            assert e in plan.outputs_of(plan.supplier[e])


# ---- who keeps a frame alive -----------------------------------------------------------------------------------------
@settings(max_examples=150, deadline=None)
@given(st.lists(st.tuples(st.sampled_from(["put", "pin", "unpin", "produce", "persist"]), st.integers(0, 3)), max_size=40))
def test_frame_snapshot_holds_exactly_what_somebody_can_still_ask_for(script):
    """FrameSnapshot's bookkeeping under arbitrary interleavings: a frame is referenced by each entry that holds it
    (once), never by an entry that is unpinned and not persistent, never released twice; equal local data shares one
    frame with one reference per entry"""
    import asyncio

    from lzy_b200.api.v1.snapshot import FrameSnapshot
    from lzy_b200.serialization.registry import LzySerializerRegistry
    from tests.test_frame_snapshot import DictStore

    store = DictStore()
    snap = FrameSnapshot(LzySerializerRegistry(), "hbm://p/", store)
    entries = [snap.create_entry(f"e{i}", int) for i in range(4)]
    pins, held, has_data, persistent = [0] * 4, [False] * 4, [False] * 4, [False] * 4

    def let_go_if_unwanted(i):  # the model of FrameSnapshot._maybe_release
        if held[i] and pins[i] == 0 and not persistent[i]:
            held[i] = False

    for action, i in script:
        e = entries[i]
        if action == "put" and not has_data[i]:
            asyncio.run(snap.put_data(e.id, i % 2))  # entries 0/2 and 1/3 carry equal data
            has_data[i] = held[i] = True
        elif action == "produce" and not has_data[i]:
            e.storage_uri = f"hbm://results/{i}"
            store.data[e.storage_uri] = i
            snap.produced(e.id)
            has_data[i] = held[i] = True
            let_go_if_unwanted(i)  # a result nobody waits for goes as soon as it exists
        elif action == "pin":
            snap.pin(e.id)
            pins[i] += 1
        elif action == "unpin" and pins[i] > 0:
            snap.unpin(e.id)
            pins[i] -= 1
            let_go_if_unwanted(i)
        elif action == "persist":
            snap.persist(e.id)
            persistent[i] = True
        else:
            continue
        holding = {}
        for j, ent in enumerate(entries):
            if held[j]:
                holding[ent._storage_uri] = holding.get(ent._storage_uri, 0) + 1
        assert store.refs == holding, (action, i, store.refs, holding)
        for uri in holding:
            assert store.exists(uri), "a frame somebody still holds was dropped"


def test_applied_env_under_contention():
    """ops of one worker share a process: while an op runs, the variables it declared have its values — whatever the
    other ops declare — and everything is restored afterwards; no schedule deadlocks (all of an op's variables are
    taken in one step)"""
    import os
    import random
    import threading

    from lzy_b200.env import applied_env

    names = ["LZYB_T_A", "LZYB_T_B", "LZYB_T_C"]
    os.environ["LZYB_T_C"] = "outer"
    errors = []

    def worker(seed):
        rnd = random.Random(seed)
        for _ in range(120):
            wanted = {n: rnd.choice(["1", "2"]) for n in names if rnd.random() < 0.6}
            with applied_env(wanted):
                for _ in range(3):
                    seen = {n: os.environ.get(n) for n in wanted}
                    if seen != wanted:
                        errors.append((wanted, seen))

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    assert not any(t.is_alive() for t in threads), "deadlock"
    assert errors == []
    assert os.environ.pop("LZYB_T_C") == "outer" and "LZYB_T_A" not in os.environ and "LZYB_T_B" not in os.environ
