#!/usr/bin/env python
"""C3 under torchrun (one worker per GPU): 1 → N-1 fan-out of a 4 GiB slot with one NCCL broadcast into
the consumers' slot frames + digest verification.  Prints one JSON line on rank 0.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/fanout_nccl.py --mib 4096
"""
import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from lzy_b200 import _native as N  # noqa: E402
from lzy_b200.fanout import broadcast_slot_nccl  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    N.init(local, 1, 16 << 20)
    n = a.mib << 20
    src = None
    if rank == 0:
        src = N.SlotHandle.create_output("c3/src", local, n)
        N.fill_pattern(local, src.device_ptr(), n, 7)
        src.commit()
    times, total = [], []
    for r in range(a.reps + 2):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        slot, ms = broadcast_slot_nccl(dist, 0, src, local, f"c3/in{r}")
        t1.record()
        torch.cuda.synchronize()
        ok = slot.info().state == N.ST_READY
        t = torch.tensor([ms, t0.elapsed_time(t1), float(ok)], device=f"cuda:{local}")
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mn = t.clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        if r >= 2:
            times.append(float(mx[0]))
            total.append(float(mx[1]))
        all_ok = bool(mn[2].item() == 1.0)
        if rank != 0:
            slot.close()
    if rank == 0:
        t = sorted(times)[len(times) // 2]
        tt = sorted(total)[len(total) // 2]
        print(json.dumps({"config": "C3", "method": "ncclBroadcast into slot frames + digest verify", "gpus": world,
                          "receivers": world - 1, "payload_mib": a.mib, "broadcast_ms": t, "time_to_ready_ms": tt,
                          "per_receiver_GBps": n / t / 1e6, "aggregate_GBps": (world - 1) * n / t / 1e6,
                          "per_receiver_GBps_incl_verify": n / tt / 1e6, "verified": all_ok}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
