#!/usr/bin/env python
"""The unsynchronised API e2e step (H2D → 8-op chain → D2H, bench.py's serial e2e leg) taken apart per iteration: CUDA events
around the three legs on the GPU, and wall time of the host-side pieces of the step (wrapped with perf_counter pairs).
A tuning aid: `python tools/e2e_trace.py [iterations] [gc|nogc]`."""
import collections
import gc
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from lzy_b200.api.v1 import B200Runtime, Lzy, materialize, op  # noqa: E402
from lzy_b200.api.v1.b200 import runtime as R  # noqa: E402
from lzy_b200.api.v1.b200 import store as S  # noqa: E402
from lzy_b200.api.v1 import snapshot as SN  # noqa: E402
from lzy_b200.utils import numa  # noqa: E402

acc = collections.Counter()


def wrap(obj, name, label):
    f = getattr(obj, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            acc[label] += time.perf_counter() - t0
    setattr(obj, name, g)


wrap(R.B200Runtime, "_exec_graph", "exec_graph")
wrap(R.B200Runtime, "_run_task", "run_task(sum)")
wrap(R.B200Runtime, "clear", "clear")
wrap(S.FrameStore, "put", "store.put")
wrap(S.FrameStore, "get", "store.get")


@op
def ident(t: torch.Tensor) -> torch.Tensor:
    return t


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    if len(sys.argv) > 2 and sys.argv[2] == "nogc":
        gc.disable()
    shape = (16384, 16384)
    rt = B200Runtime(gpus=[0], placement="single")
    lzy = Lzy(runtime=rt)
    stream = torch.cuda.Stream(0)
    torch.cuda.set_stream(stream)
    x = torch.empty(shape, device="cuda:0").normal_()
    host_in = numa.pinned_empty(shape, torch.float32, 0).copy_(x)
    host_out = numa.pinned_empty(shape, torch.float32, 0)
    dev_in = torch.empty_like(x)

    def step(inp):
        t0 = time.perf_counter()
        with lzy.workflow("p"):
            y = inp
            for _ in range(8):
                y = ident(y)
            t1 = time.perf_counter()
            out = materialize(y)
            t2 = time.perf_counter()
        t3 = time.perf_counter()
        rt.clear()
        acc["calls"] += t1 - t0
        acc["materialize"] += t2 - t1
        acc["wf exit"] += t3 - t2
        return out

    for _ in range(2):
        dev_in.copy_(host_in, non_blocking=True)
        host_out.copy_(step(dev_in), non_blocking=True)
    torch.cuda.synchronize()
    rows = []
    t_begin = time.perf_counter()
    for _ in range(iters):
        acc.clear()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        g0 = gc.get_count()
        h0 = time.perf_counter()
        e[0].record()
        dev_in.copy_(host_in, non_blocking=True)
        e[1].record()
        y = step(dev_in)
        e[2].record()
        host_out.copy_(y, non_blocking=True)
        e[3].record()
        rows.append((e, time.perf_counter() - h0, dict(acc), g0))
    torch.cuda.synchronize()
    total = time.perf_counter() - t_begin
    print(f"{1e3 * total / iters:.2f} ms per step over {iters} steps, gc {'on' if gc.isenabled() else 'off'}")
    for i, (e, host, a, g0) in enumerate(rows):
        gap = rows[i - 1][0][3].elapsed_time(e[0]) if i else 0.0
        parts = "  ".join(f"{k} {1e3 * v:.2f}" for k, v in a.items())
        print(f"  {i}: gpu gap {gap:.2f} h2d {e[0].elapsed_time(e[1]):.2f} mid {e[1].elapsed_time(e[2]):.2f} d2h {e[2].elapsed_time(e[3]):.2f}"
              f" | host {1e3 * host:.2f}: {parts} | gc counts {g0}")
    assert torch.equal(host_out, host_in)
    rt.close()


if __name__ == "__main__":
    main()
