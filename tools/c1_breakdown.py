#!/usr/bin/env python
"""Where a C1 op (1 MiB ndarray through two @ops) spends its wall time, without a profiler's distortion: the native calls
of the path are wrapped with perf_counter pairs and summed per workflow; the remainder is Python.  A tuning aid."""
import collections
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

from lzy_b200 import _host as H  # noqa: E402
from lzy_b200 import _native as N  # noqa: E402
from lzy_b200.api.v1 import B200Runtime, Lzy, materialize, op  # noqa: E402

acc = collections.Counter()
cnt = collections.Counter()


def wrap(obj, name, label=None):
    f = getattr(obj, name)
    label = label or name

    def g(*a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            acc[label] += time.perf_counter() - t0
            cnt[label] += 1
    setattr(obj, name, g)


for m in ("pack_host", "commit", "copy_to_host", "meta", "info", "wait", "close", "pack"):
    wrap(N.SlotHandle, m)
for m in ("before_execution", "after_execution", "add_slot", "close"):
    if hasattr(H.ExecutionContext, m):
        wrap(H.ExecutionContext, m, "ctx." + m)


@op
def ident(a: np.ndarray) -> np.ndarray:
    return a


def main():
    rt = B200Runtime(gpus=[0], placement="single")
    lzy = Lzy(runtime=rt)
    x = np.random.default_rng(42).standard_normal(262144, dtype=np.float32)

    def once():
        with lzy.workflow("c1"):
            return materialize(ident(ident(x)))

    for _ in range(10):
        once()
    acc.clear(), cnt.clear()
    reps = 100
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    dt = time.perf_counter() - t0
    print(f"{2 * reps / dt:.1f} ops/s, {dt / reps * 1e6:.0f} us per 2-op workflow")
    tot = 0.0
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
        print(f"  {k:22s} {cnt[k] / reps:5.1f} calls  {v / reps * 1e6:7.1f} us per workflow  ({v / cnt[k] * 1e6:6.1f} us each)")
        tot += v
    print(f"  native total {tot / reps * 1e6:.0f} us, Python remainder {(dt - tot) / reps * 1e6:.0f} us per workflow")
    rt.close()


if __name__ == "__main__":
    main()
