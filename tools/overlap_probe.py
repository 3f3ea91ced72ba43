#!/usr/bin/env python
"""What a pack (local HBM copy + digest, 0.33 ms per GiB alone) costs when it is launched on a GPU that is in the middle of
pulling 1 GiB over NVLink, and what helps: the pull as one launch or as K back-to-back sub-launches, the pack on a stream
of the same or of higher priority.  Through the C ABI; needs 2 GPUs in one process.  One JSON line per case."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from lzy_b200 import _native as N  # noqa: E402


def main():
    if torch.cuda.device_count() < 2:
        print(json.dumps({"skip": "needs 2 GPUs"}))
        return
    n = 1 << 30
    N.init(0, 2, 16 << 20)
    t1 = torch.empty(n, dtype=torch.uint8, device="cuda:1")
    N.fill_pattern(1, t1.data_ptr(), n, 7)
    x0 = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    N.fill_pattern(0, x0.data_ptr(), n, 9)
    torch.cuda.synchronize(0), torch.cuda.synchronize(1)
    src = N.SlotHandle.create_output("ov/src", 1, n)
    src.pack(t1.data_ptr(), n)
    src.commit()
    torch.cuda.set_device(0)
    s_pull = torch.cuda.Stream(0)
    s_lo, s_hi = torch.cuda.Stream(0), torch.cuda.Stream(0, priority=-1)
    engines = {"tma": N.ENGINE_TMA, "ldg": N.ENGINE_LDG}

    def run(engine, k, pack_stream, delay_ms, reps=6):
        N.set_copy_engine(engines[engine])
        pulls, packs, spans = [], [], []
        for it in range(reps + 1):
            inp = N.SlotHandle.open_input("ov/in", 0, n)
            out = N.SlotHandle.create_output("ov/out", 0, n)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            torch.cuda.synchronize(0)
            e[0].record(s_pull)
            if k:
                step = n // k
                for i in range(k):
                    inp.transfer_chunk(src, i * step, step, int(s_pull.cuda_stream))
            e[1].record(s_pull)
            if pack_stream is not None:
                if k:
                    t_end = time.perf_counter() + delay_ms * 1e-3
                    while time.perf_counter() < t_end:
                        pass
                e[2].record(pack_stream)
                out.pack(x0.data_ptr(), n, int(pack_stream.cuda_stream))
                e[3].record(pack_stream)
            torch.cuda.synchronize(0)
            if k:
                inp.finish(src, int(s_pull.cuda_stream))
                assert tuple(inp.info().digest) == tuple(src.info().digest)
            if pack_stream is not None:
                out.commit()
            if it:
                if k:
                    pulls.append(e[0].elapsed_time(e[1]))
                if pack_stream is not None:
                    packs.append(e[2].elapsed_time(e[3]))
                if k and pack_stream is not None:
                    spans.append(e[0].elapsed_time(e[3]))
            inp.close(), out.close()
        med = lambda v: round(sorted(v)[len(v) // 2], 3) if v else None  # noqa: E731
        return med(pulls), med(packs), med(spans)

    print(json.dumps({"case": "pack alone", "pack_ms": run("tma", 0, s_lo, 0)[1]}), flush=True)
    for engine in ("tma", "ldg"):
        print(json.dumps({"case": "pull alone", "engine": engine, "pull_ms": run(engine, 1, None, 0)[0],
                          "pull_8_sublaunches_ms": run(engine, 8, None, 0)[0], "pull_16_sublaunches_ms": run(engine, 16, None, 0)[0]}), flush=True)
        for k in (1, 4, 8, 16):
            for name, st in (("same priority", s_lo), ("high priority", s_hi)):
                pull, pack, span = run(engine, k, st, 0.4)
                print(json.dumps({"case": "pack launched 0.4 ms into the pull", "engine": engine, "pull_sublaunches": k, "pack_stream": name,
                                  "pull_ms": pull, "pack_wait_plus_run_ms": pack, "pull_start_to_pack_end_ms": span}), flush=True)
    # the TMA ring as a grid of several waves (CTAs retire all along the transfer) with the pack on a high-priority stream
    for waves in (2, 4, 8, 16):
        N.set_transfer_waves(waves)
        alone = run("tma", 1, None, 0)[0]
        for name, st in (("same priority", s_lo), ("high priority", s_hi)):
            pull, pack, span = run("tma", 1, st, 0.4)
            print(json.dumps({"case": "pack launched 0.4 ms into the pull", "engine": "tma", "waves": waves, "pack_stream": name,
                              "pull_alone_ms": alone, "pull_ms": pull, "pack_wait_plus_run_ms": pack, "pull_start_to_pack_end_ms": span}), flush=True)
    N.set_transfer_waves(1)
    N.set_copy_engine(N.ENGINE_AUTO)
    src.close()


if __name__ == "__main__":
    main()
