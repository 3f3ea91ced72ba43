"""The spill tier with real HBM frames: an HBM budget evicts least-recently-used frames to
files and they stay readable — by the client and as storage producers of their channels — and ``cache=True``
results written to the tier are found by a *new process* under their stable URIs, so the op is skipped
(pylzy/lzy/core/workflow.py:247-281; tests/api/v1/test_entries.py:126-330 pin the URI scheme)."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from lzy_b200 import _host as H  # noqa: E402
from lzy_b200.api.v1 import B200Runtime, Lzy, op  # noqa: E402
from lzy_b200.api.v1.utils.proxy_adapter import materialize  # noqa: E402
from lzy_b200.api.v1.b200.store import FrameStore  # noqa: E402
from lzy_b200.serialization.registry import LzySerializerRegistry  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture
def tier(native, tmp_path):
    store = FrameStore(0)
    store.configure_tier(str(tmp_path / "spill"), 0)
    yield store, LzySerializerRegistry(), tmp_path / "spill"
    store.clear("hbm://spilltest/")
    H.frame_store_set_spill(None)


def test_evicted_tensor_frame_reloads_bit_exact_with_dtype_and_shape(tier, native):
    store, ser, d = tier
    t = torch.randn(257, 1031, device="cuda:0", dtype=torch.float32)
    a = np.random.default_rng(1).integers(-5, 5, (33, 7, 5), dtype=np.int64)
    obj = {"k": [1, 2.5, "x"], "blob": b"\x00" * 1000}
    uris = ["hbm://spilltest/" + store.put("hbm://spilltest/", v, typ, ser) for v, typ in
            ((t, torch.Tensor), (a, np.ndarray), (obj, dict))]
    live0 = native.stats()["hbm_bytes_live"]
    for u in uris:
        assert store.evict(u) and not H.frame_store_resident(u) and store.exists(u)
    assert native.stats()["hbm_bytes_live"] < live0 and len(list(d.iterdir())) == 3
    h2d0 = native.stats()["h2d_bytes"]
    got_t = store.get(uris[0], torch.Tensor, ser)
    assert got_t.is_cuda and got_t.dtype == torch.float32 and torch.equal(got_t, t)
    got_a = store.get(uris[1], np.ndarray, ser)
    assert got_a.dtype == np.int64 and got_a.shape == (33, 7, 5) and np.array_equal(got_a, a)
    assert store.get(uris[2], dict, ser) == obj
    assert store.reloads == 3 and native.stats()["h2d_bytes"] - h2d0 >= t.numel() * 4 + a.nbytes
    for u in uris:  # data nobody can ask for any more takes its scratch file along
        store.drop(u)
    assert list(d.iterdir()) == []


def test_budget_evicts_least_recently_used_first(tier):
    store, ser, d = tier
    store.hbm_budget_bytes = 40 << 20
    mk = lambda i: torch.full((4 << 20,), i, dtype=torch.float32, device="cuda:0")  # 16 MiB each  # noqa: E731
    u = ["hbm://spilltest/" + store.put("hbm://spilltest/", mk(i), torch.Tensor, ser) for i in range(2)]
    assert store.evictions == 0
    store.get(u[0], torch.Tensor, ser)  # u[0] is now the most recently used
    u.append("hbm://spilltest/" + store.put("hbm://spilltest/", mk(2), torch.Tensor, ser))  # 48 MiB > 40 MiB
    assert store.evictions == 1 and not H.frame_store_resident(u[1]) and H.frame_store_resident(u[0])
    assert H.frame_store_bytes() <= 40 << 20
    assert float(store.get(u[1], torch.Tensor, ser)[123].item()) == 1.0  # comes back on demand


def test_workflow_under_an_hbm_budget_still_delivers_every_result(native, tmp_path):
    rt = B200Runtime(gpus=[0], spill_dir=str(tmp_path / "spill"), hbm_budget_bytes=24 << 20, retain="all")
    lzy = Lzy(runtime=rt)

    @op
    def scale(t: torch.Tensor, k: int) -> torch.Tensor:
        return t * k

    try:
        x = torch.arange(2 << 20, dtype=torch.float32, device="cuda:0")  # 8 MiB
        with lzy.workflow("spill-wf"):
            ys = [scale(x, k) for k in range(1, 7)]  # 6 results × 8 MiB + the argument > 24 MiB
            for k, y in enumerate(ys, 1):
                assert torch.equal(materialize(y), x * k)
        assert rt.store.evictions > 0 and rt.store.reloads > 0
    finally:
        rt.close()


CHILD = textwrap.dedent("""
    import sys, json, torch
    sys.path.insert(0, {root!r})
    from lzy_b200.api.v1 import B200Runtime, Lzy, op
    from lzy_b200.api.v1.utils.proxy_adapter import materialize
    rt = B200Runtime(gpus=[0], spill_dir={spill!r})
    lzy = Lzy(runtime=rt)

    @op(cache=True, version="1.0")
    def expensive(t: torch.Tensor, k: int) -> torch.Tensor:
        print("OP-RAN")
        return t * k + 1

    x = torch.arange(1 << 16, dtype=torch.float32, device="cuda:0")
    with lzy.workflow("cached-wf"):
        y = expensive(x, 3)
        ok = bool(torch.equal(materialize(y), x * 3 + 1))
    print(json.dumps({{"ok": ok, "ops": rt.stats["ops"], "cached": rt.stats["cached_ops"]}}))
    rt.close()
""")


def test_cached_result_survives_into_a_new_process(tmp_path):
    spill = str(tmp_path / "spill")
    outs = []
    for _ in range(2):
        r = subprocess.run([sys.executable, "-c", CHILD.format(root=str(ROOT), spill=spill)], capture_output=True, text=True,
                           timeout=300, env={**os.environ, "PYTHONPATH": str(ROOT)})
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout)
    assert "OP-RAN" in outs[0] and '"ok": true' in outs[0] and '"ops": 1' in outs[0]
    assert "OP-RAN" not in outs[1] and '"ok": true' in outs[1] and '"cached": 1' in outs[1]
