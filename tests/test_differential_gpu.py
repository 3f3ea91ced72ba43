"""Differential parity: the SAME @op graph through the oracle's restatement of the reference local runtime
(oracle/local_runtime.py: put_data → MD5 URI → FS storage → one `python startup.py` subprocess per op →
get_data, pylzy/lzy/api/v1/local/runtime.py:42-179) and through `Lzy(runtime=B200Runtime())` (HBM frames, CUDA pack /
deliver kernels, ops in-process on the GPU worker), results compared — arrays bit for bit (dtype, shape, bytes), objects
with ==, a raising op by exception type, args and the innermost traceback frame.  BASELINE configs C1, C4 (N = 569) and a
small C5 give the shapes (the workflow GPU tests asserted literals only).

The op bodies are defined inside the tests so that cloudpickle ships them by value to the oracle's subprocesses."""
import traceback
from typing import Dict

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

from lzy_b200.api.v1 import B200Runtime, Lzy, materialize, op  # noqa: E402
from oracle.local_runtime import OracleLocalRuntime  # noqa: E402


@pytest.fixture(scope="module")
def lzy():
    rt = B200Runtime(gpus=[0], placement="single")
    yield Lzy(runtime=rt)
    rt.close()


class Oracle:
    """build a graph of plain functions on the oracle runtime: call() queues, run() executes one barrier"""

    def __init__(self):
        self.rt = OracleLocalRuntime()
        self.calls = []

    def call(self, func, *args, n_outputs=1):
        c = self.rt.make_call(func, list(args), n_outputs=n_outputs)
        self.calls.append(c)
        refs = self.rt.refs(c)
        return refs[0] if n_outputs == 1 else tuple(refs)

    def run(self, *refs):
        try:
            self.rt.exec(self.calls)
            return [self.rt.materialize(r) for r in refs]
        finally:
            self.rt.close()


def same_array(a, b):
    return isinstance(a, np.ndarray) and isinstance(b, np.ndarray) and a.dtype == b.dtype and a.shape == b.shape and \
        a.tobytes() == b.tobytes()


@pytest.mark.parametrize("which", ["f32_1MiB", "i8_1e6"])
def test_c1_two_op_identity_chain(lzy, which):
    x = (np.random.default_rng(42).standard_normal(262144, dtype=np.float32) if which == "f32_1MiB"
         else np.random.default_rng(43).integers(-128, 127, 1000000, dtype=np.int8))

    def ident(a: np.ndarray) -> np.ndarray:
        return a

    o = Oracle()
    (ref,) = o.run(o.call(ident, o.call(ident, x)))
    f = op(ident)
    with lzy.workflow("diff-c1"):
        got = materialize(f(f(x)))
    assert same_array(ref, x) and same_array(got, ref)


def test_c4_bunch_through_a_train_stand_in(lzy):
    from sklearn.utils import Bunch
    rng = np.random.default_rng(0)
    bunch = Bunch(data=rng.standard_normal((569, 30)), target=rng.integers(0, 2, 569),
                  feature_names=[f"f{i}" for i in range(30)], DESCR="d" * 4096)

    def train(b: Bunch) -> np.ndarray:
        return b.data.sum(axis=0) + float(b.target.sum())

    def describe(b: Bunch) -> Dict[str, object]:
        return {"rows": int(b.data.shape[0]), "names": list(b.feature_names), "descr_len": len(b.DESCR),
                "target_dtype": str(b.target.dtype), "checksum": float(b.data[:, 0].sum())}

    o = Oracle()
    ref_w, ref_d = o.run(o.call(train, bunch), o.call(describe, bunch))
    with lzy.workflow("diff-c4"):
        w, d = op(train)(bunch), op(describe)(bunch)
        got_w, got_d = materialize(w), materialize(d)
    assert same_array(got_w, ref_w)
    assert got_d == ref_d and type(got_d) is type(ref_d)


def test_c5_small_diamond_with_mixed_payloads(lzy):
    """source → 3 branches × 2 ops → sink; every edge carries an ndarray (1 MiB) and a dict of scalars / strings"""
    def source() -> (np.ndarray, Dict[str, object]):
        return np.arange(262144, dtype=np.float32), {"hops": 0, "tag": "s", "pi": 3.25}

    def branch(t: np.ndarray, m: Dict[str, object], k: int) -> (np.ndarray, Dict[str, object]):
        return t * np.float32(k) + np.float32(1), {"hops": m["hops"] + 1, "tag": m["tag"] + str(k), "pi": m["pi"]}

    def sink(a: np.ndarray, b: np.ndarray, c: np.ndarray, ma: Dict[str, object], mb: Dict[str, object],
             mc: Dict[str, object]) -> (np.ndarray, Dict[str, object]):
        return a + b + c, {"hops": ma["hops"] + mb["hops"] + mc["hops"], "tags": [ma["tag"], mb["tag"], mc["tag"]]}

    o = Oracle()
    t, m = o.call(source, n_outputs=2)
    ends = []
    for k in (2, 3, 5):
        t1, m1 = o.call(branch, t, m, k, n_outputs=2)
        ends.append(o.call(branch, t1, m1, k + 1, n_outputs=2))
    rt_, rm_ = o.call(sink, ends[0][0], ends[1][0], ends[2][0], ends[0][1], ends[1][1], ends[2][1], n_outputs=2)
    ref_t, ref_m = o.run(rt_, rm_)

    src, br, snk = op(source), op(branch), op(sink)
    with lzy.workflow("diff-c5"):
        t, m = src()
        ends = []
        for k in (2, 3, 5):
            t1, m1 = br(t, m, k)
            ends.append(br(t1, m1, k + 1))
        gt, gm = snk(ends[0][0], ends[1][0], ends[2][0], ends[0][1], ends[1][1], ends[2][1])
        got_t, got_m = materialize(gt), materialize(gm)
    assert same_array(got_t, ref_t)
    assert got_m == ref_m


def test_a_raising_op_surfaces_identically(lzy):
    """type, args and the frame that raised (file, line, function, source text) — pylzy/lzy/api/v1/startup.py:137-170,
    local/runtime.py:165-173: the caller sees the op's own exception with the op's traceback"""
    def double(a: np.ndarray) -> np.ndarray:
        return a * 2

    def boom(a: np.ndarray) -> np.ndarray:
        if a.sum() > 0:
            raise ValueError("bad total", int(a.sum()), {"where": "boom"})
        return a

    x = np.arange(10, dtype=np.int64)

    def innermost(exc):
        frames = traceback.extract_tb(exc.__traceback__)
        ours = [f for f in frames if f.name == "boom"]
        assert ours, [f.name for f in frames]
        f = ours[-1]
        return (f.filename, f.lineno, f.name, f.line)

    o = Oracle()
    with pytest.raises(ValueError) as ref:
        o.run(o.call(double, o.call(boom, o.call(double, x))))
    with pytest.raises(ValueError) as got:
        with lzy.workflow("diff-exc"):
            materialize(op(double)(op(boom)(op(double)(x))))
    assert type(got.value) is type(ref.value) and got.value.args == ref.value.args == ("bad total", 90, {"where": "boom"})
    assert innermost(got.value) == innermost(ref.value)
