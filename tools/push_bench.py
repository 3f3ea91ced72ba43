#!/usr/bin/env python
"""GPU→GPU slot delivery through the product C ABI (lzyb_transfer_chunk + lzyb_input_finish) for every carrier —
push kernel + landed scan, TMA-ring pull, LDG pull, copy engine + trailing scans — at several sizes, with one
direction busy and with both (each GPU source and sink at once, the ring bench's situation).  CUDA-event timed on
the sink streams, max over the two directions.  One JSON line per point.  Needs 2 GPUs in one process."""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from lzy_b200 import _native as N  # noqa: E402

ENGINES = {"push": N.ENGINE_PUSH, "tma": N.ENGINE_TMA, "ldg": N.ENGINE_LDG, "dma": N.ENGINE_DMA}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, nargs="+", default=[1, 8, 32, 128, 1024])
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", nargs="*", default=None, help="carriers to run (default: all): push tma ldg dma")
    ap.add_argument("--dirs", type=int, nargs="*", default=[1, 2])
    args = ap.parse_args()
    if torch.cuda.device_count() < 2:
        print(json.dumps({"skip": "needs 2 GPUs"}))
        return
    N.init(0, 2, 16 << 20)
    streams = [torch.cuda.Stream(device=g) for g in (0, 1)]
    sps = [int(s.cuda_stream) for s in streams]
    for mib in args.mib:
        n = mib << 20
        outs = []
        for g in (0, 1):
            t = torch.empty(n, dtype=torch.uint8, device=f"cuda:{g}")
            N.fill_pattern(g, t.data_ptr(), n, 5 + g)
            torch.cuda.synchronize(g)
            o = N.SlotHandle.create_output(f"pb/out{g}", g, n)
            o.pack(t.data_ptr(), n)
            o.commit()
            outs.append(o)
            del t
        for name, eng in ENGINES.items():
            if args.only and name not in args.only:
                continue
            N.set_copy_engine(eng)
            for dirs in args.dirs:
                sinks = [1] if dirs == 1 else [1, 0]  # sink GPU g pulls from out[1 - g]

                def once():
                    inps = []
                    for g in sinks:
                        inp = N.SlotHandle.open_input(f"pb/in{g}", g, n)
                        inp.transfer_chunk(outs[1 - g], 0, n, sps[g])
                        inps.append((g, inp))
                    for g, inp in inps:
                        inp.finish(outs[1 - g], sps[g])
                    for g, inp in inps:
                        assert tuple(inp.info().digest) == tuple(outs[1 - g].info().digest)
                        inp.close()

                for _ in range(3):
                    once()
                for g in (0, 1):
                    torch.cuda.synchronize(g)
                ev = {g: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for g in sinks}
                for g in sinks:
                    ev[g][0].record(streams[g])
                for _ in range(args.reps):
                    once()
                for g in sinks:
                    ev[g][1].record(streams[g])
                for g in (0, 1):
                    torch.cuda.synchronize(g)
                ms = max(ev[g][0].elapsed_time(ev[g][1]) for g in sinks) / args.reps
                print(json.dumps({"carrier": name, "payload_MiB": mib, "directions": dirs, "ms_per_transfer": round(ms, 4),
                                  "GBps_per_dir": round(n / ms / 1e6, 1),
                                  "includes": "open_input + transfer + finish (host waits for the verdict) + close, per transfer"}),
                      flush=True)
        N.set_copy_engine(N.ENGINE_AUTO)
        for o in outs:
            o.close()


if __name__ == "__main__":
    main()
