#!/usr/bin/env python
"""bench.py — the slot/channel data plane on BASELINE.json's configs[1]:
"linear 8-@op chain, 1 GiB fp32 tensor per edge, 1→8 GPUs peer-copy".

One *step* = one pass of the chain: ingest x into a slot, then for each of the 8 ops the consumer's input slot
receives the producer's output slot (verified against the producer's digest), the op body runs (identity) and its
result is packed + framed + digested into a fresh output slot.  Every hop is therefore one slot→slot delivery of
`payload` bytes; a step delivers 8 × payload per GPU.

  N = 1   all slots on one GPU.  The consumer is the producer's only reader on the same GPU, so the delivery is an
          ownership hand-over of the frame (lzyb_slot_set_movable; FileInputBackend.toOutput in the reference) and
          the op's result is packed into a fresh frame: 9 full-payload launches per step (1 ingest + 8 packs).
  N > 1   launched by torchrun, one worker process per GPU.  The N chains are staggered round a ring: at every hop
          rank r's output is PUSHED into rank r+1's input frame over NVLink: the TMA-ring kernel runs on r's GPU with the
          frame (mapped through CUDA IPC) as its destination, folds the digest from its shared-memory stages, verifies it
          against the digest r committed and writes header, verdict and arrival word; r+1's GPU runs a one-thread waiter
          on that word (push_kernels.cuh).  Units are independent (no
          data-path collective; only 256-byte slot descriptors cross a shared-memory mailbox) and per-GPU work is
          fixed → "scaling": "weak".

`value`   device-resident throughput, inputs already in HBM when the timed region starts.
`e2e`     the same chain driven from HOST memory through the public API (`Lzy(runtime=B200Runtime()).workflow()`):
          pinned input → H2D → 8 @ops → result → D2H into pinned memory, copies inside the timed region.  At N > 1
          ONE runtime spans all N GPUs (rank 0's process; op k of chain j on GPU (j + k) mod N, N chains per step).
`parity`  before anything is timed: the chain's final frame against the regenerated source tensor (torch.equal) once
          per carrier, and a corrupted producer caught as a peer fault.  The run fails otherwise.
`--impl reference`  the reference's own CPU local runtime for this path (oracle/local_runtime.py — a restatement;
          the reference's JVM/pylzy stack cannot be installed offline) on the same config; each step
          is a bounded sample of it (REF_SAMPLE_OPS ops of the chain at the full 1 GiB payload).
`--config c1|c3|c4|c5`  the other BASELINE configs alone; the default run appends them under `extra.configs`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_OPS = 8
METRIC = "slot_to_slot_GBps"
UNIT = "GB/s"
REF_SAMPLE_OPS = int(os.environ.get("LZYB_REF_SAMPLE_OPS", "2"))
WORKLOAD = "linear 8-@op chain, 1 GiB fp32 tensor per edge"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c1", "c3", "c4", "c5"])
    ap.add_argument("--payload-mib", type=int, default=1024)
    ap.add_argument("--engine", "--carrier", dest="engine", default="auto", choices=["auto", "push", "ldg", "tma", "dma"],
                    help="how a frame crosses GPUs (N > 1): push = producer-issued push kernel + landed scan (auto); "
                         "ldg / tma / dma = consumer pull through a CUDA-IPC import.  N = 1: ldg / tma force a copying pull")
    ap.add_argument("--no-move", action="store_true", help="N = 1: copy on delivery even for a sole consumer (round-1 behaviour)")
    ap.add_argument("--blocking-commit", action="store_true",
                    help="N = 1: every op waits for its pack and digest before the next delivery (lzyb_slot_commit as round 1 used it)")
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C1/C3/C4/C5 records")
    return ap.parse_args()


def workload_config(payload: int, world: int) -> dict:
    """the SAME dict in both arms: it names the workload, not how an arm carries it (ours adds `carrier_detail`)"""
    mib = payload >> 20
    return {"workload": WORKLOAD if mib == 1024 else f"linear 8-@op chain, {mib} MiB fp32 tensor per edge",
            "payload_bytes": payload, "ops": N_OPS, "gpus": world,
            "placement": "1 GPU" if world == 1 else f"ring of {world} GPUs, one worker process per GPU, chain j enters on GPU j",
            "l2": "inputs (1 GiB/edge) larger than the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's restatement of the reference local runtime
# ------------------------------------------------------------------------------------------------
def _ident(a):
    return a


def cpu_chain_seconds(payload_bytes: int, n_ops: int) -> float:
    import numpy as np
    from oracle.local_runtime import run_chain
    x = np.random.default_rng(1234).standard_normal(payload_bytes // 4, dtype=np.float32)
    t0 = time.perf_counter()
    y = run_chain([_ident] * n_ops, x)
    dt = time.perf_counter() - t0
    assert y.shape == x.shape and y[0] == x[0] and y[-1] == x[-1]
    return dt


def cpu_sample_text(payload: int, n_ops: int) -> str:
    return (f"{n_ops} consecutive @ops of the {N_OPS}-op identity chain at the full {payload >> 20} MiB per edge (put_data of x, "
            f"{n_ops} ops each in its own subprocess reading and writing the edge through the FS storage, get_data of the result): "
            f"a bounded sample of the workload — the chain's fixed ends are amortised over {n_ops} ops instead of {N_OPS}, which "
            f"UNDERSTATES the port's throughput by ≈ 10–15 % (measured: 256 MiB, 1 op 4.6 s, 2 ops 7.9 s)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    payload = args.payload_mib << 20
    n_ops = max(1, min(N_OPS, REF_SAMPLE_OPS))
    for _ in range(args.warmup):
        cpu_chain_seconds(payload, n_ops)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_chain_seconds(payload, n_ops)
    val = n_ops * payload * args.steps / t / 1e9
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(payload, args.gpus),
        "ops_per_sec": n_ops * args.steps / t,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": cpu_sample_text(payload, n_ops),
                         "sample_ops_per_step": n_ops, "host_cores": os.cpu_count()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler (pynvml; same fields as the profiling recipe's nvidia-smi line)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Ring:
    """everything one worker of the chain bench needs: device, streams, the mailbox to its neighbours"""

    SINK, OUT = 0, 1  # mailbox lanes

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        from lzy_b200 import _native as N
        self.torch, self.dist, self.N = torch, dist, N
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.multi = self.world > 1
        self.stream_order = False  # N = 1: commits complete in stream order (set in run_ours)
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device; the data plane has no CPU path")
        self.dev = self.local
        torch.cuda.set_device(self.dev)
        self.host_group = None
        self.box = None
        if self.multi:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.dev}"))
            # host-side barriers for the legs in which one rank must have the GPUs to itself: an NCCL barrier would
            # leave a spinning kernel on every waiting GPU
            self.host_group = dist.new_group(backend="gloo")
            from lzy_b200.ring import ShmMailbox
            self.box = ShmMailbox.rendezvous(_GlooFacade(dist, self.host_group), self.world, self.rank, lanes=2)
            self.pids = [None] * self.world  # rank 0 later waits for the other ranks' processes to be gone (wait_until_alone)
            dist.all_gather_object(self.pids, os.getpid(), group=self.host_group)
        N.init(self.dev, 1, 64 << 20)
        if args.ctas_per_sm:
            N.set_grid(args.ctas_per_sm)
        self.payload = args.payload_mib << 20
        n_elem = self.payload // 4
        side = int(round(n_elem ** 0.5))
        self.shape = (side, side) if side * side == n_elem else (n_elem,)
        self.stream = torch.cuda.Stream(self.dev)       # a real (non-legacy) stream: every launch and every event goes here
        self.push_stream = torch.cuda.Stream(self.dev)  # this GPU as a source: the push towards the successor
        torch.cuda.set_stream(self.stream)
        self.sp, self.sp_push = int(self.stream.cuda_stream), int(self.push_stream.cuda_stream)
        assert self.sp != 0 and self.sp_push != 0
        self.x = self.source_tensor(self.rank)
        from lzy_b200.utils import numa
        self.host_in = numa.pinned_empty(self.shape, torch.float32, self.dev)  # pinned, on the socket this GPU hangs off
        self.host_in.copy_(self.x)
        self.host_out = numa.pinned_empty(self.shape, torch.float32, self.dev)
        torch.cuda.synchronize()
        self.kern_events = []  # (start, end, kind) for every data-moving launch of ours inside the timed region
        self.step_seq = 0

    def source_tensor(self, rank: int):
        """the chain input of `rank` (torch.Generator(seed).normal_); regenerated anywhere for the parity check"""
        torch = self.torch
        gen = torch.Generator(device=f"cuda:{self.dev}")
        gen.manual_seed(1234 + rank)
        x = torch.empty(self.shape, dtype=torch.float32, device=f"cuda:{self.dev}")
        x.normal_(generator=gen)
        return x

    def barrier(self):
        if self.multi:
            self.dist.barrier()

    def host_barrier(self):
        if self.multi:
            self.dist.barrier(group=self.host_group)

    def timed(self, kind, fn, record):
        if not record:
            return fn()
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        r = fn()
        e1.record(self.stream)
        self.kern_events.append((e0, e1, kind))
        return r

    # -- one delivery: producer slot `prev` (mine) → my new input slot; returns the READY input slot -----------------
    def deliver(self, prev, name: str, carrier: str, record: bool, move: bool):
        N, torch = self.N, self.torch
        inp = N.SlotHandle.open_input(name, self.dev, self.payload)
        if not self.multi:
            if move:
                prev.set_movable(True)  # this chain is the slot's only consumer: what the channel manager's peer table says
            moved = self.timed("deliver", lambda: inp.transfer_chunk(prev, 0, self.payload, self.sp), record)
            assert moved == self.payload
            inp.finish(prev, self.sp)
            return inp
        from lzy_b200.ring import predecessor, successor
        pred, succ = predecessor(self.rank, self.world), successor(self.rank, self.world)
        if carrier == "push":
            self.box.post(self.SINK, bytes(inp.sink_export()))
            self.box.post(self.OUT, bytes(prev.export()))
            sink = N.IpcDesc.from_buffer_copy(self.box.read(succ, self.SINK))
            src = N.IpcDesc.from_buffer_copy(self.box.read(pred, self.OUT))
            # my GPU as a source first: my committed output goes into my successor's frame.  (The first push into a frame
            # maps it — cudaIpcOpenMemHandle synchronises this device — so it must not find my own waiter spinning.)
            prev.push(sink, 0, 0, self.sp_push)
            # ... and as a sink: one thread waits for my predecessor's arrival word, then hands over verdict and digest
            got = self.timed("deliver", lambda: inp.expect_push(src, 0, 0, self.sp), record)
            assert got == self.payload
            inp.finish_expect(src.digest, self.sp)
            return inp
        self.box.post(self.OUT, bytes(prev.export()))
        peer = N.SlotHandle.import_peer(N.IpcDesc.from_buffer_copy(self.box.read(pred, self.OUT)), self.dev)
        try:
            moved = self.timed("deliver", lambda: inp.transfer_chunk(peer, 0, self.payload, self.sp), record)
            assert moved == self.payload
            inp.finish(peer, self.sp)
        finally:
            peer.close()
        return inp

    def chain_step(self, record: bool, from_host: bool, carrier: str, move: bool, corrupt_at=None):
        """returns (final output slot, every slot to close).  `corrupt_at`: flip one byte of the ingest slot after its
        commit — the first delivery must then fail as a peer fault (returns the exception instead of the final slot)."""
        from lzy_b200.devmem import view_tensor
        N, torch = self.N, self.torch
        self.step_seq += 1
        sid = self.step_seq
        live = []
        ingest = N.SlotHandle.create_output(f"s{sid}/ingest", self.dev, self.payload)
        ingest.set_meta(N.DT_F32, self.shape)
        live.append(ingest)
        if from_host:
            ingest.pack_host(self.host_in.data_ptr(), self.payload, self.sp)
        else:
            self.timed("pack", lambda: ingest.pack(self.x.data_ptr(), self.payload, self.sp), record)
        if self.stream_order and corrupt_at is None:
            ingest.commit_async(self.sp)
        else:
            ingest.commit(self.sp)
        if corrupt_at is not None:
            ingest.fault(N.FAULT_CORRUPT_ONCE, corrupt_at)
        prev = ingest
        for k in range(N_OPS):
            try:
                inp = self.deliver(prev, f"s{sid}/op{k}/in", carrier, record, move and corrupt_at is None)
            except N.ReadException as e:
                if corrupt_at is None:
                    raise
                return e, live
            live.append(inp)
            if corrupt_at is not None:
                return None, live  # the corruption went through unnoticed
            # the op body: identity on a zero-copy view of the delivered slot
            arg = view_tensor(inp.device_ptr(), torch.float32, self.shape, self.dev)
            res = arg
            out = N.SlotHandle.create_output(f"s{sid}/op{k}/out", self.dev, self.payload)
            out.set_meta(N.DT_F32, self.shape)
            live.append(out)
            self.timed("pack", lambda: out.pack(res.data_ptr(), self.payload, self.sp), record)
            if self.stream_order:
                out.commit_async(self.sp)  # READY in stream order: the next op's launches queue behind this pack
            else:
                out.commit(self.sp)
            prev = out
        if from_host:
            prev.copy_to_host(self.host_out.data_ptr(), self.payload, 0, self.sp)
        if self.stream_order:
            prev.commit(self.sp)  # the step's one wait: the chain's last pack has finished and its digest is published
        return prev, live

    def close_all(self, live):
        if self.multi:
            # nobody recycles a frame a neighbour may still be reading or writing
            self.torch.cuda.synchronize()
            self.barrier()
        for sl in live:
            sl.close()

    def run(self, steps, record, from_host, carrier, move):
        for _ in range(steps):
            _last, live = self.chain_step(record, from_host, carrier, move)
            self.close_all(live)

    def measure(self, steps, from_host, carrier, move):
        torch = self.torch
        self.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        self.run(steps, not from_host, from_host, carrier, move)
        e1.record(self.stream)
        torch.cuda.synchronize()
        self.barrier()
        ms = e0.elapsed_time(e1)
        return self.max_over_ranks(ms)

    def max_over_ranks(self, v: float) -> float:
        if not self.multi:
            return v
        t = self.torch.tensor([v], device=f"cuda:{self.dev}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, ok: bool) -> bool:
        if not self.multi:
            return ok
        t = self.torch.tensor([1 if ok else 0], device=f"cuda:{self.dev}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    # -- parity, before anything is timed --------------------------------------------------------------------------------
    def parity(self, carriers, move_default):
        """bit-exact delivery through the whole chain, once per carrier: the final frame against the regenerated
        source tensor of the rank the chain started on, and its digest against the ingest slot's; then a corrupted
        producer must be caught as a peer fault by every carrier that verifies on delivery"""
        from lzy_b200.devmem import view_tensor
        N, torch = self.N, self.torch
        src_rank = (self.rank - N_OPS) % self.world
        want = self.x if src_rank == self.rank else self.source_tensor(src_rank)
        want_digest = N.digest_device(self.dev, want.data_ptr(), self.payload, self.sp)
        checked, ok_all, caught_all = [], True, True
        for carrier in carriers:
            N.set_copy_engine({"ldg": N.ENGINE_LDG, "tma": N.ENGINE_TMA, "dma": N.ENGINE_DMA}.get(carrier, N.ENGINE_AUTO))
            move = move_default and carrier == "move"
            last, live = self.chain_step(False, False, carrier, move)
            got = view_tensor(last.device_ptr(), torch.float32, self.shape, self.dev)
            ok = bool(torch.equal(got, want)) and tuple(last.info().digest) == tuple(want_digest)
            self.close_all(live)
            ok = self.all_true(ok)
            # one wrong byte in the producer's frame after its commit (SlotsTest's failing peer, at the data level)
            res, live = self.chain_step(False, False, carrier, False, corrupt_at=(self.payload // 3) | 1)
            caught = isinstance(res, N.ReadException)
            self.close_all(live)
            caught = self.all_true(caught)
            checked.append({"carrier": carrier, "bit_exact": ok, "corrupted_producer_is_a_peer_fault": caught})
            ok_all, caught_all = ok_all and ok, caught_all and caught
        N.set_copy_engine(N.ENGINE_AUTO)
        if src_rank != self.rank:
            del want
        block = {"engines": [c["carrier"] for c in checked], "bit_exact": ok_all, "corruption_detected": caught_all,
                 "per_carrier": checked,
                 "how": ("torch.equal(final frame of the 8-op chain, regenerated input of rank (r − 8) mod N) and equal digests on every rank; "
                         "then LZYB_FAULT_CORRUPT_ONCE on the producer's committed frame must surface as LZYB_E_PEER at the consumer")}
        if not (ok_all and caught_all):
            raise SystemExit(f"bench.py: parity check failed, nothing was timed: {json.dumps(block)}")
        return block


class _GlooFacade:
    """ShmMailbox.rendezvous needs broadcast_object_list + barrier on a HOST group"""

    def __init__(self, dist, group):
        self._d, self._g = dist, group

    def broadcast_object_list(self, objs, src=0):
        self._d.broadcast_object_list(objs, src=src, group=self._g)

    def barrier(self):
        self._d.barrier(group=self._g)


def run_ours(args):
    import torch

    ring = Ring(args)
    N, multi, world, rank, dev, payload = ring.N, ring.multi, ring.world, ring.rank, ring.dev, ring.payload

    if args.config != "c2":
        return run_single_config(args, ring)

    # which carrier the timed chain uses, and which ones the parity check also exercises
    if multi:
        carrier = "push" if args.engine == "auto" else args.engine
        parity_carriers = ["push", "tma", "ldg", "dma"]
        move = False
    else:
        carrier = "move" if (args.engine == "auto" and not args.no_move) else ("ldg" if args.engine in ("auto", "push", "dma") else args.engine)
        parity_carriers = ["move", "ldg", "tma"]
        move = carrier == "move"
    # N = 1: commits complete in stream order (lzyb_set_lazy_commit — B200Runtime's default): an op's hand-over and pack are queued
    # behind its predecessor's pack instead of waiting for it on the host; the step waits once, for the last digest.  The ring
    # (N > 1) exports every committed frame to another process, which needs the digest: it commits as before.
    ring.stream_order = (not multi) and not args.blocking_commit
    N.set_lazy_commit(ring.stream_order)
    parity = ring.parity(parity_carriers, True)
    N.set_copy_engine({"ldg": N.ENGINE_LDG, "tma": N.ENGINE_TMA, "dma": N.ENGINE_DMA}.get(carrier, N.ENGINE_AUTO))

    ring.run(args.warmup, False, False, carrier, move)
    N.lib.lzyb_stats_reset()
    ring.kern_events.clear()
    sampler = ClockSampler(dev)
    sampler.start()
    ms_total = ring.measure(args.steps, False, carrier, move)
    clocks = sampler.stop()
    st = N.stats()
    launches = st["kernel_launches"]

    # live roofline numbers for the dominant kernel: CUDA events around every data-moving launch in the timed region
    per_kind = {}
    for e0, e1, kind in ring.kern_events:
        per_kind.setdefault(kind, []).append(e0.elapsed_time(e1))
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    packs = per_kind.get("pack", [])
    delivers = per_kind.get("deliver", [])
    timed_ms = sum(packs) + (sum(delivers) if carrier != "move" else 0.0)
    if not multi:
        avg_ms = sum(packs) / len(packs)
        algo = 2 * payload  # bytes read + bytes written per launch (2·S per pack on one GPU)
        achieved = algo / (avg_ms * 1e-3) / 1e9
        full_launches = (len(packs) + (len(delivers) if carrier != "move" else 0)) / args.steps
        # a same-GPU edge costs 2·S of HBM traffic; the step has 8 edges + the ingest pack
        step_algo = (N_OPS + 1) * 2 * payload
        step_moved = full_launches * 2 * payload
        roof = {"bound": "hbm", "kernel": "stream_digest_kernel<16,8,true> (LDG.128/STG.128 + digest, 2 tiles per CTA): pack",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                "peak_source": peak_src, "avg_launch_ms": avg_ms, "launches_timed": len(packs),
                "kernel_share_of_step": timed_ms / ms_total,
                "full_payload_launches_per_step": full_launches, "frames_moved_per_step": st.get("frames_moved", 0) / args.steps,
                "step_hbm_bytes_over_algorithmic": step_moved / step_algo,
                "step_level": {"algorithmic_bytes": step_algo, "achieved": step_algo / (ms_total / args.steps * 1e-3) / 1e9,
                               "frac": step_algo / (ms_total / args.steps * 1e-3) / 1e9 / hbm_peak,
                               "note": "9 launches × 2·S per step (ingest + 8 packs; deliveries are ownership hand-overs) ÷ ms_per_step"}}
        traffic_file = ROOT / "profiles" / "traffic.json"
        if traffic_file.exists():
            try:
                roof["traffic"] = json.loads(traffic_file.read_text()).get("dram_bytes_per_launch")
            except Exception:
                pass
    else:
        pavg = sum(delivers) / len(delivers)
        achieved = payload / (pavg * 1e-3) / 1e9  # S bytes cross NVLink per delivery
        names = {"push": "tma_stream_digest_kernel<3,32768,true> launched on the SOURCE GPU (cp.async.bulk global → shared → peer global through the "
                         "CUDA-IPC mapping, digest folded from shared memory); timed on the SINK stream = arrival_wait_kernel, i.e. ready-to-receive → verified frame",
                 "tma": "tma_stream_digest_kernel<3,32768,true> (consumer pull, TMA bulk ring)",
                 "ldg": "stream_digest_kernel<16,8,true> (consumer pull, LDG.128 over NVLink)",
                 "dma": "copy engine in geometric sub-chunks + stream_digest_kernel<16,8,false> trailing scans (consumer pull)"}
        roof = {"bound": "nvlink", "kernel": names[carrier], "achieved": achieved, "peak": 770.0, "unit": "GB/s",
                "frac": achieved / 770.0, "traffic": None,
                "peak_source": "measured peer copy 770 GB/s/dir (B200_PROFILING.md); nominal 900",
                "frac_of_nominal_900": achieved / 900.0, "avg_launch_ms": pavg, "launches_timed": len(delivers),
                "dma_copies": st.get("dma_copies", 0), "push_launches": st.get("push_launches", 0),
                "kernel_share_of_step": timed_ms / ms_total,
                "sm_push_ceiling": {"GBps": 714, "why": "every SM-issued NVLink write is a 128-byte packet: 1.1875 wire bytes per payload byte "
                                                         "(profiles/r02_peer_variants_ncu.csv); the copy engine's 772 needs larger packets"}}

    value = world * N_OPS * payload * args.steps / (ms_total * 1e-3) / 1e9

    # ---- e2e through the C ABI from host buffers (every rank its own chain) ------------------------------------------
    e2e_cabi = None
    if not args.no_e2e:
        ring.run(1, False, True, carrier, move)
        e2e_steps = max(3, min(args.steps, 10))
        ms_e2e = ring.measure(e2e_steps, True, carrier, move)
        ok = torch.equal(ring.host_out, ring.host_in) if not multi else True
        if multi:  # after 8 hops host_out holds the input of rank (r − 8) mod N
            want = ring.source_tensor((rank - N_OPS) % world).cpu()
            ok = bool(torch.equal(ring.host_out, want))
        if not ring.all_true(ok):
            raise SystemExit("bench.py: the host-buffer chain delivered different bytes")
        e2e_cabi = {"value": world * N_OPS * payload * e2e_steps / (ms_e2e * 1e-3) / 1e9, "unit": UNIT,
                    "h2d_bytes_per_step": world * payload, "d2h_bytes_per_step": world * payload, "ms_per_step": ms_e2e / e2e_steps,
                    "steps": e2e_steps, "bit_exact": True,
                    "api": "C ABI: lzyb_slot_pack_host → chain → lzyb_copy_to_host (pinned host buffers), one chain per GPU"}

    # ---- extras that need every rank (C3 through NCCL) -----------------------------------------------------------------
    extras = []
    if multi and world >= 4 and not args.no_extras:
        import bench_configs
        rec = bench_configs.c3_nccl(ring.dist, rank, world, dev)
        if rec:
            extras.append(rec)

    # ---- from here on rank 0 works alone (one runtime over all GPUs); the others wait on the host ----------------------
    torch.cuda.synchronize()
    ring.host_barrier()
    if rank != 0:
        # Leave: the rest is one process driving every GPU.  A second process with a CUDA context on a GPU — even an idle one —
        # makes the launches of the first one there slower (time-slicing between contexts): the API step measured
        # 60 ms with rank 0 alone on the box and 74 ms with rank 1 parked on GPU 1.
        if ring.box is not None:
            ring.box.close()
        N.shutdown()
        # ... and leave for good: a teardown that waits for rank 0 (which is about to measure for a while) would keep this
        # process and its CUDA context on the GPU — then it is cut short
        t = threading.Thread(target=ring.dist.destroy_process_group, daemon=True)
        t.start()
        t.join(5.0)
        if t.is_alive():
            sys.stdout.flush()
            os._exit(0)
        return
    api, e2e, cpu = None, e2e_cabi, None
    if rank == 0:
        gpus = list(range(world))
        if multi:
            N.shutdown()  # this process managed GPU 0 only; the runtime below manages all N
            waited_s = wait_until_alone([p for r, p in enumerate(ring.pids) if r != rank])
        if not args.no_e2e:
            api_steps = max(3, min(args.steps, 20))  # as many as the headline when the driver asks for 20: one stalled step weighs 5 %
            rec = api_bench(dev, gpus, payload, ring.shape, api_steps)
            api = {"device_resident": {"value": world * N_OPS * payload / (rec["ms_dev"] * 1e-3) / 1e9, "unit": UNIT,
                                       "ms_per_step": rec["ms_dev"], "ops_per_sec": world * N_OPS / (rec["ms_dev"] * 1e-3)},
                   "steps": api_steps, "chains_per_step": world, "gpus": world, "stats": rec["stats"],
                   "call": "with Lzy(runtime=B200Runtime(gpus=all)).workflow(): ys = [f8(...f1(x_j)) for j in range(N)]; materialize each"}
            if multi:
                api["other_ranks_gone_after_s"] = round(waited_s, 2)  # rank 0 measures alone on the box (wait_until_alone)
            e2e = {"value": world * N_OPS * payload / (rec["ms_e2e"] * 1e-3) / 1e9, "unit": UNIT,
                   "h2d_bytes_per_step": world * payload, "d2h_bytes_per_step": world * payload, "ms_per_step": rec["ms_e2e"],
                   "steps": api_steps, "ops_per_sec": world * N_OPS / (rec["ms_e2e"] * 1e-3), "bit_exact": True,
                   "phases_ms": rec["phases"], "gpu_timeline_ms": rec["timeline"],
                   "api": ("Lzy(runtime=B200Runtime(gpus=[0..N-1])).workflow(): x_j.to(cuda:j) from pinned host → 8 @op chain, op k of chain j on GPU "
                           "(j + k) mod N → y_j.cpu() into pinned host; ONE runtime in one process spans all N GPUs"),
                   "c_abi_only": e2e_cabi}
            if rec.get("ms_duplex") is not None:
                e2e["duplex"] = {"value": world * N_OPS * payload / (rec["ms_duplex"] * 1e-3) / 1e9, "unit": UNIT,
                                 "ms_per_step": rec["ms_duplex"],
                                 "how": "result D2H of step i on a second stream, overlapping the input H2D of step i+1"}
        if not args.no_extras:
            import bench_configs
            for fn in ((lambda: bench_configs.c1(dev, with_cpu=not multi)),
                       (lambda: bench_configs.c4(dev, rows_list=(569, 1_000_000) if not multi else (569,), with_cpu=not multi)),
                       (lambda: bench_configs.c5(gpus)),
                       (lambda: bench_configs.c3_in_process(gpus) if world >= 4 else None),
                       (lambda: bench_configs.worker_processes_chain(gpus) if world >= 2 else None)):
                try:
                    r = fn()
                    if isinstance(r, list):
                        extras.extend(r)
                    elif r:
                        extras.append(r)
                except Exception as e:  # an extra never takes the headline down; it says what happened
                    extras.append({"config": "extra", "error": f"{type(e).__name__}: {e}"})
        if not multi and not args.no_cpu_baseline:
            n_ops = max(1, min(N_OPS, REF_SAMPLE_OPS))
            dt = cpu_chain_seconds(payload, n_ops)
            cpu = {"value": n_ops * payload / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": cpu_sample_text(payload, n_ops) + f"; one pass, {dt:.1f} s", "sample_ops_per_step": n_ops,
                   "host_cores": os.cpu_count()}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": workload_config(payload, world),
            "carrier_detail": {"carrier": carrier, "engine_flag": args.engine,
                               "commit": ("stream order (lzyb_set_lazy_commit): launches of op k+1 queue behind the pack of op k; one host wait per "
                                          "step, on the last digest" if ring.stream_order else "blocking: the host waits for every pack's digest"),
                               "what": ("ownership hand-over of the frame to its sole same-GPU consumer + pack of the op result" if carrier == "move" else
                                        "TMA-ring push kernel on the producer's GPU (digest folded, header + arrival word written into the consumer's frame)" if carrier == "push" else
                                        f"consumer pull ({carrier})")},
            "ops_per_sec": world * N_OPS * args.steps / (ms_total * 1e-3),
            "roofline": roof, "parity": parity, "cpu_baseline": cpu, "e2e": e2e, "api": api, "clocks": clocks, "gpu_launches": launches,
            "launch_breakdown": {k: {"n": len(v), "avg_ms": sum(v) / len(v), "min_ms": min(v), "median_ms": sorted(v)[len(v) // 2], "max_ms": max(v)}
                                 for k, v in per_kind.items()},
            "per_edge_GBps": per_edge_gbps(per_kind.get("deliver" if multi else "pack"), payload, multi),
            "extra": {"configs": extras},
        }), flush=True)
    if ring.box is not None:
        ring.box.close()
    if multi:
        ring.dist.destroy_process_group()


def per_edge_gbps(ms, payload: int, multi: bool):
    """per-edge GB/s, min / median / max over the deliveries of the timed region.  N > 1: the NVLink
    delivery of S bytes.  N = 1: an edge is a hand-over (no bytes move) + the pack of the consumer's result, so the edge
    costs that pack (S bytes of payload; 2·S of HBM traffic)."""
    if not ms:
        return None
    v = sorted(ms)
    gbps = lambda t: payload / (t * 1e-3) / 1e9  # noqa: E731
    return {"min": gbps(v[-1]), "median": gbps(v[len(v) // 2]), "max": gbps(v[0]), "edges": len(v),
            "what": "NVLink delivery, S bytes per edge" if multi else "the pack that follows the hand-over, S payload bytes per edge"}


def wait_until_alone(pids, timeout_s: float = 20.0) -> float:
    """the other ranks are exiting: wait until their processes are gone (no /proc entry, or a zombie waiting for torchrun to
    reap it — its CUDA contexts are released either way), at most `timeout_s`.  Returns the seconds waited.  The process ids
    come from the ranks themselves, so they are in this process's pid namespace (NVML's process list need not be)."""
    def gone(pid: int) -> bool:
        try:
            with open(f"/proc/{pid}/stat") as f:
                return f.read().rsplit(")", 1)[1].split()[0] in ("Z", "X")
        except OSError:
            return True

    t0 = time.time()
    while time.time() - t0 < timeout_s and not all(gone(p) for p in pids):
        time.sleep(0.05)
    if all(gone(p) for p in pids):
        time.sleep(0.2)  # the driver's teardown of a context that died with its process
    return time.time() - t0


def api_bench(dev0: int, gpus, payload: int, shape, steps: int):
    """The chain through the public API — the call a user makes — on ONE runtime spanning `gpus`:

        with Lzy(runtime=B200Runtime(gpus=gpus)).workflow("bench"):
            ys = [f8(f7(... f1(x_j))) for j in range(len(gpus))]

    one chain per GPU (chain j enters on GPU j, op k runs on GPU (j + k) mod N), so every GPU hosts 8 ops per step
    and, for N > 1, every edge crosses NVLink.
    device: the x_j are CUDA tensors already, results stay on the device.
    e2e:    the x_j sit in pinned host memory; each step copies them into the caller's device tensors (H2D), runs the
            chains and copies every y_j into pinned host memory (D2H)."""
    import numpy as np
    import torch

    from lzy_b200 import _native as N
    from lzy_b200.api.v1 import B200Runtime, Lzy, materialize, op
    from lzy_b200.utils import numa
    globals().update(torch=torch, np=np)  # annotations are strings here (PEP 563) and are evaluated in module globals
    n = len(gpus)

    def make_ops(j):
        fs = []
        for k in range(N_OPS):
            g = gpus[(j + k) % n]

            @op
            def ident(t: torch.Tensor) -> torch.Tensor:
                return t
            fs.append(ident.with_provisioning(gpu_index=g) if n > 1 else ident)
        return fs

    chains = [make_ops(j) for j in range(n)]
    rt = B200Runtime(gpus=list(gpus), placement="single" if n == 1 else "round_robin")
    lzy = Lzy(runtime=rt)
    streams = [torch.cuda.current_stream(g) if g == dev0 else torch.cuda.Stream(g) for g in gpus]
    xs, host_in, host_out, dev_in = [], [], [], []
    for j, g in enumerate(gpus):
        x = torch.empty(shape, dtype=torch.float32, device=f"cuda:{g}").normal_()
        xs.append(x)
        dev_in.append(torch.empty_like(x))  # the caller's own staging tensor for the H2D copy (no allocation per step)
        host_in.append(numa.pinned_empty(shape, torch.float32, g).copy_(x))  # each GPU's host buffers on its own socket
        host_out.append(numa.pinned_empty(shape, torch.float32, g))
    d2h_streams = [torch.cuda.Stream(g) for g in gpus]

    def step(inputs):
        with lzy.workflow("bench"):
            ys = []
            for j, inp in enumerate(inputs):
                y = inp
                for f in chains[j]:
                    y = f(y)
                ys.append(y)
            outs = [materialize(y) for y in ys]
        rt.clear()  # release the HBM frames of this step
        return outs

    def sync_all():
        for g in gpus:
            torch.cuda.synchronize(g)

    phases = {"h2d": 0.0, "chain": 0.0, "d2h": 0.0}
    marks = []  # chain 0 of every serial e2e step: CUDA events around its three legs, on the stream that carries them

    def timed(k, from_host, duplex=False, collect=False, timeline=False):
        sync_all()
        t0 = time.perf_counter()
        outs = None
        for _ in range(k):
            if from_host:
                p0 = time.perf_counter()
                inputs = []
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timeline else None
                for j, g in enumerate(gpus):
                    with torch.cuda.stream(streams[j]):
                        if ev and j == 0:
                            ev[0].record()
                        inputs.append(dev_in[j].copy_(host_in[j], non_blocking=True))
                        if ev and j == 0:
                            ev[1].record()
                if collect:
                    sync_all()
                p1 = time.perf_counter()
                outs = step(inputs)
                if collect:
                    sync_all()
                p2 = time.perf_counter()
                for j, y in enumerate(outs):
                    ydev = y.device.index
                    if duplex:
                        d2h_streams[gpus.index(ydev)].wait_stream(torch.cuda.current_stream(ydev))
                        with torch.cuda.stream(d2h_streams[gpus.index(ydev)]):
                            host_out[j].copy_(y, non_blocking=True)
                        y.record_stream(d2h_streams[gpus.index(ydev)])
                    else:
                        # the D2H rides the stream the H2D of this GPU rides (never the legacy default stream, which would
                        # fence every blocking stream of the device), ordered behind the stream the result was handed over on
                        streams[gpus.index(ydev)].wait_stream(torch.cuda.current_stream(ydev))
                        with torch.cuda.stream(streams[gpus.index(ydev)]):
                            if ev and j == 0:
                                ev[2].record()
                            host_out[j].copy_(y, non_blocking=True)
                            if ev and j == 0:
                                ev[3].record()
                if ev:
                    marks.append(ev)
                if collect:
                    sync_all()
                    p3 = time.perf_counter()
                    phases["h2d"] += (p1 - p0) * 1e3 / k
                    phases["chain"] += (p2 - p1) * 1e3 / k
                    phases["d2h"] += (p3 - p2) * 1e3 / k
            else:
                outs = step(xs)
        sync_all()
        return (time.perf_counter() - t0) * 1e3, outs

    _, outs = timed(2, False)
    for j in range(n):
        assert torch.equal(outs[j].to(xs[j].device), xs[j]), "API chain delivered different bytes"
    N.lib.lzyb_stats_reset()
    ms_dev, _ = timed(steps, False)
    st = N.stats()
    stats = {k: st[k] // steps for k in ("kernel_launches", "push_launches", "frames_moved", "dma_copies", "pack_launches", "pull_launches")}
    timed(3, True)  # warm-up of the host leg: staging tensors, pinned blocks and the frame pool have seen the pattern
    import gc
    gc_log = []  # (generation, ms) of every collection inside the serial e2e leg

    def gc_cb(phase, info, _t=[0.0]):
        if phase == "start":
            _t[0] = time.perf_counter()
        else:
            gc_log.append((info["generation"], round(1e3 * (time.perf_counter() - _t[0]), 2)))
    gc.callbacks.append(gc_cb)
    ms_e2e, _ = timed(steps, True, timeline=True)
    gc.callbacks.remove(gc_cb)
    tl = {"h2d": 0.0, "h2d_end_to_d2h_start": 0.0, "d2h": 0.0, "d2h_end_to_next_h2d": 0.0}
    one_gpu = (N_OPS - 1) % n == 0  # chain 0 ends on the GPU it entered on: its four events share a device
    for i, ev in enumerate(marks):
        tl["h2d"] += ev[0].elapsed_time(ev[1]) / len(marks)
        tl["d2h"] += ev[2].elapsed_time(ev[3]) / len(marks)
        if one_gpu:
            tl["h2d_end_to_d2h_start"] += ev[1].elapsed_time(ev[2]) / len(marks)
            if i:
                tl["d2h_end_to_next_h2d"] += marks[i - 1][3].elapsed_time(ev[0]) / (len(marks) - 1)
    mids = [round(ev[1].elapsed_time(ev[2]), 2) for ev in marks] if one_gpu else None
    if not one_gpu:  # events of different devices cannot be subtracted
        tl["h2d_end_to_d2h_start"] = tl["d2h_end_to_next_h2d"] = float("nan")
    for j in range(n):
        assert torch.equal(host_out[j], host_in[j]), "API e2e chain delivered different bytes"
    timed(2, True, collect=True)  # the same step with a device sync between its three legs: where the time goes
    ms_duplex = None
    try:  # same work per step, the two PCIe directions of neighbouring steps overlapped
        for h in host_out:
            h.zero_()
        timed(1, True, duplex=True)
        ms, _ = timed(steps, True, duplex=True)
        if all(torch.equal(host_out[j], host_in[j]) for j in range(n)):
            ms_duplex = ms / steps
    except Exception as e:  # noqa: BLE001 — the serial number stands
        print(f"[bench] duplex e2e leg skipped: {type(e).__name__}: {e}", file=sys.stderr, flush=True)
    rt.close()
    return {"ms_dev": ms_dev / steps, "ms_e2e": ms_e2e / steps, "ms_duplex": ms_duplex, "stats": stats,
            "phases": {k: round(v, 3) for k, v in phases.items()}, "timeline": {**{k: round(v, 3) for k, v in tl.items() if v == v}, "h2d_end_to_d2h_start_per_step": mids,
                         "gc_collections_ms": [g for g in gc_log if g[1] >= 0.5]}}


def run_single_config(args, ring):
    """`--config c1|c3|c4|c5`: one of the other BASELINE configs alone, same JSON contract"""
    import bench_configs
    world, rank, dev = ring.world, ring.rank, ring.dev
    rec = None
    t0 = time.perf_counter()
    if args.config == "c3":
        if world >= 2:
            rec = bench_configs.c3_nccl(ring.dist, rank, world, dev)
        else:
            rec = {"config": "C3", "skipped": "a fan-out needs consumers on other GPUs: run with --gpus ≥ 2 (policy: NCCL for ≥ 3 receivers)"}
    else:
        ring.host_barrier()
        if rank == 0:
            if ring.multi:
                ring.N.shutdown()
            gpus = list(range(world))
            rec = {"c1": lambda: bench_configs.c1(dev), "c4": lambda: bench_configs.c4(dev), "c5": lambda: bench_configs.c5(gpus)}[args.config]()
        ring.host_barrier()
    if rank == 0:
        value, unit, metric = None, "ops/s", "op_chain_ops_per_sec"
        if args.config == "c1":
            value = rec["f32_1MiB"]["ops_per_sec"]
        elif args.config == "c4":
            value = rec["runs"][0]["ops_per_sec"]
        elif args.config == "c5":
            value = rec["prefetch_on"]["ops_per_sec"]
        elif args.config == "c3" and "per_receiver_GBps" in rec:
            value, unit, metric = rec["aggregate_GBps"], "GB/s", "fanout_aggregate_GBps"
        print(json.dumps({"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                          "data": "synthetic", "config": {"workload": rec.get("workload", rec.get("config")), "which": args.config},
                          "record": rec, "wall_s": time.perf_counter() - t0}), flush=True)
    if ring.box is not None:
        ring.box.close()
    if ring.multi:
        ring.dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
