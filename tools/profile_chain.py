#!/usr/bin/env python
"""Wall-clock breakdown of the 8-op 1 GiB chain through the @op API (device-resident input)."""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from lzy_b200 import _native as N  # noqa: E402
from lzy_b200.api.v1 import B200Runtime, Lzy, materialize, op  # noqa: E402


@op
def ident(t: torch.Tensor) -> torch.Tensor:
    return t


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1  # GPUs: n chains, op k of chain j on GPU (j + k) mod n
    gpus = list(range(n))
    rt = B200Runtime(gpus=gpus, placement="single" if n == 1 else "round_robin")
    lzy = Lzy(runtime=rt)
    xs = [torch.empty(16384, 16384, device=f"cuda:{g}").normal_() for g in gpus]
    chains = [[ident.with_provisioning(gpu_index=gpus[(j + k) % n]) if n > 1 else ident for k in range(8)] for j in range(n)]

    def sync():
        for g in gpus:
            torch.cuda.synchronize(g)

    timelines = []
    for it in range(int(os.environ.get("ITERS", "6"))):
        sync()
        t0 = time.perf_counter()
        with lzy.workflow("p"):
            t1 = time.perf_counter()
            ys = []
            for j in range(n):
                y = xs[j]
                for f in chains[j]:
                    y = f(y)
                ys.append(y)
            t2 = time.perf_counter()
            out = [materialize(y) for y in ys]
            t2b = time.perf_counter()
            sync()
            t3 = time.perf_counter()
        t4 = time.perf_counter()
        rt.clear()
        t5 = time.perf_counter()
        st = N.stats()
        if rt.trace is not None:
            agg = {}
            for _t, ph, ms, _at in rt.trace:
                agg[ph] = agg.get(ph, 0.0) + ms
            print("   ", {k: round(v, 2) for k, v in agg.items()})
            if n > 1 and it >= 2:  # every step as a timeline: when each op's phases ended, ms after the workflow was entered
                per = {}
                for t, ph, ms, at in rt.trace:
                    per.setdefault(t, []).append(f"{ph[:6]} {ms:.2f}@{1e3 * (at - t0):.2f}")
                lines = [f"      task {t:2d}: " + "  ".join(per[t]) for t in sorted(per)]
                lines.append(f"      barrier entered @{1e3 * (t2 - t0):.2f}, materialize returned @{1e3 * (t2b - t0):.2f}")
                timelines.append((t3 - t2, it, lines))
        if n > 1:
            log = rt.transfer_log(clear=True)
            big = [r for r in log if r.get("bytes", 0) >= 1 << 29]
            ms = sorted(r["ns"] / 1e6 for r in big)
            if big and it == 5:
                for k in ("ns_setup", "ns_issue", "ns_wait", "ns_complete"):
                    v = sorted(r[k] / 1e6 for r in big)
                    print(f"      {k}: median {v[len(v) // 2]:.3f} max {v[-1]:.3f} ms")
            if ms:
                print(f"    {len(ms)} transfers: min {ms[0]:.2f} median {ms[len(ms) // 2]:.2f} max {ms[-1]:.2f} ms;"
                      f" materialize returned at {1e3 * (t2b - t2):.2f}, device idle at {1e3 * (t3 - t2):.2f} ms")
            rt.trace.clear()
        print(f"iter {it}: enter {1e3*(t1-t0):.2f} calls {1e3*(t2-t1):.2f} exec+materialize {1e3*(t3-t2):.2f} exit {1e3*(t4-t3):.2f} "
              f"clear {1e3*(t5-t4):.2f} ms | pooled {st['hbm_bytes_pooled']>>20} MiB live {st['hbm_bytes_live']>>20} MiB launches {st['kernel_launches']}")
    if timelines:
        timelines.sort()
        for label, (dur, it, lines) in (("fastest", timelines[0]), ("slowest", timelines[-1])):
            print(f"    {label} step: iter {it}, exec+materialize {1e3 * dur:.2f} ms")
            print("\n".join(lines))
    if len(sys.argv) > 2 and sys.argv[2] == "kernels":
        # one more step under the CUPTI tracer (torch.profiler): the kernels of every GPU on one time axis
        from torch.profiler import ProfilerActivity, profile
        sync()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            with lzy.workflow("p"):
                ys = []
                for j in range(n):
                    y = xs[j]
                    for f in chains[j]:
                        y = f(y)
                    ys.append(y)
                out = [materialize(y) for y in ys]
                sync()
        evs = [e for e in prof.events() if e.device_type.name == "CUDA" and e.device_time_total > 20]
        evs.sort(key=lambda e: e.time_range.start)
        base = evs[0].time_range.start if evs else 0
        for e in evs:
            print(f"      gpu{e.device_index} {1e-3 * (e.time_range.start - base):8.3f} ms +{1e-3 * e.device_time_total:6.3f}  {e.name[:60]}")
        rt.clear()
    rt.close()


if __name__ == "__main__":
    main()
