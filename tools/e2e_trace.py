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
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 1  # GPUs: n chains, op k of chain j on GPU (j + k) mod n
    gpus = list(range(n))
    shape = (16384, 16384)
    rt = B200Runtime(gpus=gpus, placement="single" if n == 1 else "round_robin")
    lzy = Lzy(runtime=rt)
    streams = [torch.cuda.Stream(g) for g in gpus]
    torch.cuda.set_stream(streams[0])
    xs = [torch.empty(shape, device=f"cuda:{g}").normal_() for g in gpus]
    host_in = [numa.pinned_empty(shape, torch.float32, g).copy_(xs[g]) for g in gpus]
    host_out = [numa.pinned_empty(shape, torch.float32, g) for g in gpus]
    dev_in = [torch.empty_like(x) for x in xs]
    chains = [[ident.with_provisioning(gpu_index=gpus[(j + k) % n]) if n > 1 else ident for k in range(8)] for j in range(n)]

    def step(inputs):
        t0 = time.perf_counter()
        with lzy.workflow("p"):
            ys = []
            for j, inp in enumerate(inputs):
                y = inp
                for f in chains[j]:
                    y = f(y)
                ys.append(y)
            t1 = time.perf_counter()
            outs = [materialize(y) for y in ys]
            t2 = time.perf_counter()
        t3 = time.perf_counter()
        rt.clear()
        acc["calls"] += t1 - t0
        acc["materialize"] += t2 - t1
        acc["wf exit"] += t3 - t2
        return outs

    def sync():
        for g in gpus:
            torch.cuda.synchronize(g)

    def one(ev=None):
        inputs = []
        for j, g in enumerate(gpus):
            with torch.cuda.stream(streams[j]):
                if ev:
                    ev[j][0].record()
                inputs.append(dev_in[j].copy_(host_in[j], non_blocking=True))
                if ev:
                    ev[j][1].record()
        outs = step(inputs)
        for j, y in enumerate(outs):
            d = y.device.index
            streams[d].wait_stream(torch.cuda.current_stream(d))
            with torch.cuda.stream(streams[d]):
                if ev:
                    ev[d][2].record()
                host_out[j].copy_(y, non_blocking=True)
                if ev:
                    ev[d][3].record()

    for _ in range(2):
        one()
    sync()
    # one reference event per GPU, recorded back to back: the time axes of the devices agree to within the issue jitter
    refs = []
    for j, g in enumerate(gpus):
        with torch.cuda.stream(streams[j]):
            r = torch.cuda.Event(enable_timing=True)
            r.record()
            refs.append(r)
    rows = []
    t_begin = time.perf_counter()
    for _ in range(iters):
        acc.clear()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in gpus]
        h0 = time.perf_counter()
        one(ev)
        rows.append((ev, time.perf_counter() - h0, dict(acc), 1e3 * (h0 - t_begin)))
    sync()
    total = time.perf_counter() - t_begin
    print(f"{1e3 * total / iters:.2f} ms per step over {iters} steps on {n} GPU(s), gc {'on' if gc.isenabled() else 'off'}")
    for i, (ev, host, a, h_at) in enumerate(rows):
        parts = "  ".join(f"{k} {1e3 * v:.2f}" for k, v in a.items())
        print(f"  step {i} issued @{h_at:.2f} | host {1e3 * host:.2f}: {parts}")
        for j in range(n):
            t = [refs[j].elapsed_time(e) for e in ev[j]]
            print(f"      gpu{j}: h2d {t[0]:.2f} → {t[1]:.2f}   d2h {t[2]:.2f} → {t[3]:.2f}")
    for j in range(n):
        assert torch.equal(host_out[j], host_in[j])
    rt.close()


if __name__ == "__main__":
    main()
