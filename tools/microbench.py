#!/usr/bin/env python
"""Kernel-level sweep: pack (same-GPU copy+digest) and pull (same GPU, and peer GPU when ≥ 2 are
visible in this process) for every engine / ring shape / grid, CUDA-event timed, inputs larger than
L2.  Prints one JSON line per point.  Not the bench: a tuning aid whose output goes to profiles/."""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from lzy_b200 import _native as N  # noqa: E402


def time_op(fn, stream, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, nargs="+", default=[1024])
    ap.add_argument("--peer", action="store_true")
    args = ap.parse_args()
    ngpu = torch.cuda.device_count()
    N.init(0, 0, 16 << 20)
    dst_gpu = 1 if (args.peer and ngpu > 1) else 0
    torch.cuda.set_device(dst_gpu)
    stream = torch.cuda.Stream(dst_gpu)
    torch.cuda.set_stream(stream)
    sp = int(stream.cuda_stream)
    variants = [("ldg", None, c) for c in (2, 4, 8)] + [("tma", i, 0) for i in range(6)]
    for mib in args.mib:
        n = mib << 20
        src_t = torch.empty(n, dtype=torch.uint8, device="cuda:0")
        N.fill_pattern(0, src_t.data_ptr(), n, 3)
        torch.cuda.synchronize(0)
        src = N.SlotHandle.create_output("mb/src", 0, n)
        src.pack(src_t.data_ptr(), n)
        src.commit()
        for eng, cfg, ctas in variants:
            N.set_copy_engine(N.ENGINE_TMA if eng == "tma" else N.ENGINE_LDG)
            if cfg is not None:
                N.set_tma_config(cfg)
            if ctas:
                N.set_grid(ctas)

            def pull():
                inp = N.SlotHandle.open_input("mb/in", dst_gpu, n)
                inp.transfer_chunk(src, 0, n, sp)
                inp.finish(src, sp)
                inp.close()

            def pack():
                out = N.SlotHandle.create_output("mb/out", 0, n)
                out.pack(src_t.data_ptr(), n, sp)
                out.commit(sp)
                out.close()

            ms_pull = time_op(pull, stream)
            rec = {"mib": mib, "engine": eng, "tma_cfg": cfg, "ctas_per_sm": ctas or None, "dst_gpu": dst_gpu,
                   "pull_ms": round(ms_pull, 4), "pull_payload_GBps": round(n / ms_pull / 1e6, 1)}
            if dst_gpu == 0:
                torch.cuda.set_device(0)
                ms_pack = time_op(pack, stream)
                rec.update(pack_ms=round(ms_pack, 4), pack_payload_GBps=round(n / ms_pack / 1e6, 1),
                           hbm_traffic_GBps=round(2 * n / min(ms_pack, ms_pull) / 1e6, 1))
            print(json.dumps(rec), flush=True)
        src.close()
        N.set_copy_engine(N.ENGINE_AUTO)
        N.set_tma_config(1)
        N.set_grid(-1000)


if __name__ == "__main__":
    main()
