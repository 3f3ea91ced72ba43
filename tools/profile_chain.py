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
    rt = B200Runtime(gpus=[0], placement="single")
    lzy = Lzy(runtime=rt)
    x = torch.empty(16384, 16384, device="cuda:0").normal_()
    for it in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with lzy.workflow("p"):
            t1 = time.perf_counter()
            y = x
            for _ in range(8):
                y = ident(y)
            t2 = time.perf_counter()
            out = materialize(y)
            torch.cuda.synchronize()
            t3 = time.perf_counter()
        t4 = time.perf_counter()
        rt.clear()
        t5 = time.perf_counter()
        st = N.stats()
        if rt.trace is not None:
            agg = {}
            for _t, ph, ms in rt.trace:
                agg[ph] = agg.get(ph, 0.0) + ms
            print("   ", {k: round(v, 2) for k, v in agg.items()})
            rt.trace.clear()
        print(f"iter {it}: enter {1e3*(t1-t0):.2f} calls {1e3*(t2-t1):.2f} exec+materialize {1e3*(t3-t2):.2f} exit {1e3*(t4-t3):.2f} "
              f"clear {1e3*(t5-t4):.2f} ms | pooled {st['hbm_bytes_pooled']>>20} MiB live {st['hbm_bytes_live']>>20} MiB launches {st['kernel_launches']}")
    rt.close()


if __name__ == "__main__":
    main()
