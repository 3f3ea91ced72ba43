#!/usr/bin/env python
"""BASELINE.json configs C2–C5 on the in-process multi-GPU runtime (one worker pinned per GPU of this
box).  The driver's bench line is bench.py (configs[1]); this records the other configs for DESIGN.md /
profiles/.  One JSON line per measurement.

  C2  linear 8-@op chain, 1 GiB fp32 per edge, op k on GPU k mod G (through the @op API)
  C3  1→G fan-out of a 4 GiB slot: chunk-pipelined pull chain (lzyb_broadcast) vs G-1 concurrent unicasts
      vs NCCL broadcast is measured by tools/fanout_nccl.py under torchrun
  C4  synthetic sklearn Bunch (N rows × 30 f64 + i64 target + strings) through a `train` stand-in op
  C5  16-op diamond (source → 7 branches × 2 → sink), 256 MiB tensor + ~64 KiB dict per edge, prefetch on/off
"""
import argparse
import json
import sys
import threading
import time
from pathlib import Path
from typing import Dict

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from lzy_b200 import _native as N  # noqa: E402
from lzy_b200.api.v1 import B200Runtime, Lzy, materialize, op  # noqa: E402

G = torch.cuda.device_count()


def emit(**kw):
    print(json.dumps(kw), flush=True)


@op
def ident(t: torch.Tensor) -> torch.Tensor:
    return t


def c2(mib: int, reps: int):
    rt = B200Runtime(placement="round_robin")
    lzy = Lzy(runtime=rt)
    n = (mib << 20) // 4
    x = torch.empty(n, device="cuda:0").normal_()
    times = []
    for r in range(reps + 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with lzy.workflow("c2"):
            y = x
            for _ in range(8):
                y = ident(y)
            out = materialize(y)
        for g in range(G):
            torch.cuda.synchronize(g)
        dt = time.perf_counter() - t0
        if r >= 2:
            times.append(dt)
        ok = torch.equal(out.to("cuda:0"), x)
        rt.clear()
    t = sorted(times)[len(times) // 2]
    emit(config="C2", gpus=G, payload_mib=mib, ops=8, wall_ms=t * 1e3, ops_per_sec=8 / t, edge_GBps=8 * (mib << 20) / t / 1e9,
         bit_exact=bool(ok), placement="op k on GPU k mod G")
    rt.close()


def c3(mib: int, reps: int):
    N.init(0, 0, 16 << 20)
    n = mib << 20
    src_t = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    N.fill_pattern(0, src_t.data_ptr(), n, 7)
    torch.cuda.synchronize(0)
    src = N.SlotHandle.create_output("c3/src", 0, n)
    src.pack(src_t.data_ptr(), n)
    dg = src.commit()
    consumers = [g for g in range(G) if g != 0] or [0]

    def run_broadcast(chunk):
        dsts = [N.SlotHandle.open_input(f"c3/d{g}", g, n) for g in consumers]
        t0 = time.perf_counter()
        N.broadcast(src, dsts, chunk)
        dt = time.perf_counter() - t0
        ok = all(tuple(d.info().digest) == dg for d in dsts)
        for d in dsts:
            d.close()
        return dt, ok

    def run_unicasts():
        dsts = [N.SlotHandle.open_input(f"c3/u{g}", g, n) for g in consumers]
        streams = [torch.cuda.Stream(g) for g in consumers]
        t0 = time.perf_counter()
        for d, s in zip(dsts, streams):
            d.transfer_chunk(src, 0, n, int(s.cuda_stream))
        for d in dsts:
            d.finish(src)
        dt = time.perf_counter() - t0
        ok = all(tuple(d.info().digest) == dg for d in dsts)
        for d in dsts:
            d.close()
        return dt, ok

    for name, fn in [("device-chained kernel (1 launch per GPU)", lambda: run_broadcast(0)), ("pull-chain 16MiB", lambda: run_broadcast(16 << 20)), ("pull-chain 32MiB", lambda: run_broadcast(32 << 20)),
                     ("pull-chain 64MiB", lambda: run_broadcast(64 << 20)), ("pull-chain 128MiB", lambda: run_broadcast(128 << 20)),
                     ("pull-chain 256MiB", lambda: run_broadcast(256 << 20)), ("concurrent unicasts", run_unicasts)]:
        ts = []
        for r in range(reps + 1):
            dt, ok = fn()
            if r:
                ts.append(dt)
        t = sorted(ts)[len(ts) // 2]
        emit(config="C3", gpus=G, payload_mib=mib, receivers=len(consumers), method=name, time_to_last_ms=t * 1e3,
             per_receiver_GBps=n / t / 1e9, aggregate_GBps=len(consumers) * n / t / 1e9, verified=bool(ok))
    src.close()


def c4(rows: int, reps: int):
    from sklearn.utils import Bunch
    rng = np.random.default_rng(0)
    bunch = Bunch(data=rng.standard_normal((rows, 30)), target=rng.integers(0, 2, rows),
                  feature_names=[f"f{i}" for i in range(30)], DESCR="d" * 4096)

    @op
    def train(b: Bunch) -> np.ndarray:  # stand-in for CatBoost fit: a column reduction
        return b.data.sum(axis=0) + float(b.target.sum())

    rt = B200Runtime(placement="single")
    lzy = Lzy(runtime=rt)
    ts = []
    for r in range(reps + 1):
        t0 = time.perf_counter()
        with lzy.workflow("c4"):
            res = materialize(train(bunch))
        dt = time.perf_counter() - t0
        if r:
            ts.append(dt)
        rt.clear()
    ok = np.allclose(res, bunch.data.sum(axis=0) + float(bunch.target.sum()))
    t = sorted(ts)[len(ts) // 2]
    nbytes = bunch.data.nbytes + bunch.target.nbytes
    emit(config="C4", rows=rows, array_bytes=nbytes, wall_ms=t * 1e3, ops_per_sec=1 / t, GBps=nbytes / t / 1e9, correct=bool(ok))
    rt.close()


def c5(mib: int, reps: int):
    meta = {f"k{i}": ("v" * 40, i, float(i)) for i in range(700)}  # ≈ 64 KiB pickled

    @op
    def source() -> (torch.Tensor, Dict[str, tuple]):
        return torch.ones((mib << 20) // 4, device="cuda"), meta

    @op
    def branch(t: torch.Tensor, m: Dict[str, tuple]) -> (torch.Tensor, Dict[str, tuple]):
        t = t.mul_(1.0)
        for _ in range(12):  # busy work of about one transfer time
            t = t * 1.0
        return t, m

    @op
    def sink(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, d: torch.Tensor, e: torch.Tensor, f: torch.Tensor,
             g: torch.Tensor) -> float:
        dev = a.device
        return float(sum(float(x.to(dev)[0].item()) for x in (a, b, c, d, e, f, g)))

    for prefetch in (True, False):
        rt = B200Runtime(placement="round_robin", prefetch=prefetch)
        lzy = Lzy(runtime=rt)
        ts = []
        for r in range(reps + 1):
            for gdev in range(G):
                torch.cuda.synchronize(gdev)
            t0 = time.perf_counter()
            with lzy.workflow("c5"):
                t, m = source()
                ends = []
                for _ in range(7):
                    t1, m1 = branch(t, m)
                    t2, m2 = branch(t1, m1)
                    ends.append(t2)
                res = materialize(sink(*ends))
            dt = time.perf_counter() - t0
            if r:
                ts.append(dt)
            rt.clear()
        tmed = sorted(ts)[len(ts) // 2]
        emit(config="C5", gpus=G, payload_mib=mib, ops=16, prefetch=prefetch, wall_ms=tmed * 1e3, ops_per_sec=16 / tmed,
             correct=bool(res == 7.0))
        rt.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=["C2", "C3", "C4", "C5"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--c3-mib", type=int, default=4096)
    a = ap.parse_args()
    if "C2" in a.only:
        c2(1024, a.reps)
    if "C3" in a.only:
        c3(a.c3_mib, a.reps)
    if "C4" in a.only:
        c4(569, a.reps)
        c4(1_000_000, max(2, a.reps // 2))
    if "C5" in a.only:
        c5(256, a.reps)
