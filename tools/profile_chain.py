#!/usr/bin/env python
"""Wall-clock breakdown of the 8-op 1 GiB chain through the @op API (device-resident input)."""
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

    for it in range(6):
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
            if it == 5 and n > 1:  # the last step as a timeline: when each op's phases ended, ms after the first mark
                base = min(at - ms * 1e-3 for _t, _ph, ms, at in rt.trace)
                per = {}
                for t, ph, ms, at in rt.trace:
                    per.setdefault(t, []).append(f"{ph[:6]} {ms:.2f}@{1e3 * (at - base):.2f}")
                for t in sorted(per):
                    print(f"      task {t:2d}: " + "  ".join(per[t]))
            rt.trace.clear()
        print(f"iter {it}: enter {1e3*(t1-t0):.2f} calls {1e3*(t2-t1):.2f} exec+materialize {1e3*(t3-t2):.2f} exit {1e3*(t4-t3):.2f} "
              f"clear {1e3*(t5-t4):.2f} ms | pooled {st['hbm_bytes_pooled']>>20} MiB live {st['hbm_bytes_live']>>20} MiB launches {st['kernel_launches']}")
    rt.close()


if __name__ == "__main__":
    main()
