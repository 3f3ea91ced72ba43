#!/usr/bin/env python
"""cProfile of the @op path on small payloads (where host overhead dominates): C1's two-op chain."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

from lzy_b200.api.v1 import B200Runtime, Lzy, materialize, op  # noqa: E402


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

    for _ in range(5):
        once()
    t0 = time.perf_counter()
    for _ in range(50):
        once()
    dt = time.perf_counter() - t0
    print(f"{100 / dt:.1f} ops/s, {dt / 50 * 1e3:.3f} ms per 2-op workflow")
    import threading
    prof = cProfile.Profile()
    threading.setprofile(lambda *a: None)
    prof.enable()
    for _ in range(50):
        once()
    prof.disable()
    if len(sys.argv) > 2:
        prof.dump_stats(sys.argv[2])  # raw stats for pstats on another box
    pstats.Stats(prof).sort_stats(sys.argv[1] if len(sys.argv) > 1 else "cumulative").print_stats(45)


if __name__ == "__main__":
    main()
