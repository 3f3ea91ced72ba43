"""BASELINE.json configs C1, C3, C4, C5 — the workloads next to the headline C2 — as functions that return one
record each, so that `bench.py` can attach them to its JSON line (`extra.configs`) and run any of them alone
(`bench.py --config c1|c3|c4|c5`).  SURVEY.md §8(d) defines the shapes.

  C1  two-@op identity chain over a 1 MiB float32 ndarray: ops/s through the public API on the B200 runtime, next to the
      same chain through the CPU restatement of the reference's local runtime (one subprocess per op), same host
  C3  1→R fan-out of a 4 GiB slot (R = GPUs − 1 consumers on distinct GPUs): one ncclBroadcast straight into the
      consumers' slot frames (+ digest verification) vs the device-chained pull kernel vs R concurrent unicasts
  C4  synthetic sklearn Bunch through a `train` stand-in op, arrays as out-of-band parts of one frame
  C5  16-op diamond (source → 7 branches × 2 ops → sink), 256 MiB tensor + ~64 KiB dict per edge, prefetch on / off,
      overlap measured from CUDA events

The CPU figures come from `oracle/` (test infrastructure: only the timed baseline legs may call it).
"""
import time
from typing import Any, Dict, List, Optional


def _median(xs: List[float]) -> float:
    s = sorted(xs)
    return s[len(s) // 2]


# ------------------------------------------------------------------------------------------------------------------
# C1
# ------------------------------------------------------------------------------------------------------------------
def c1(dev: int = 0, reps: int = 20, cpu_reps: int = 3, with_cpu: bool = True) -> Dict[str, Any]:
    import numpy as np

    from lzy_b200.api.v1 import B200Runtime, Lzy, materialize, op

    @op
    def ident(a: np.ndarray) -> np.ndarray:
        return a

    out: Dict[str, Any] = {"config": "C1", "workload": "two-@op identity chain, 1 MiB float32 ndarray (262144 elements)"}
    rt = B200Runtime(gpus=[dev], placement="single")
    lzy = Lzy(runtime=rt)
    for name, x in (("f32_1MiB", np.random.default_rng(42).standard_normal(262144, dtype=np.float32)),
                    ("i8_1e6", np.random.default_rng(43).integers(-128, 127, 1000000, dtype=np.int8))):
        def once():
            with lzy.workflow("c1"):
                return materialize(ident(ident(x)))

        for _ in range(2):
            y = once()
        assert np.array_equal(y, x) and y.dtype == x.dtype
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        dt = time.perf_counter() - t0
        out[name] = {"ops_per_sec": 2 * reps / dt, "ms_per_op": 1e3 * dt / (2 * reps), "reps": reps, "bit_exact": True}
    rt.close()
    if with_cpu:
        from oracle.local_runtime import run_chain
        x = np.random.default_rng(42).standard_normal(262144, dtype=np.float32)
        run_chain([_ident, _ident], x)
        ts = []
        for _ in range(cpu_reps):
            t0 = time.perf_counter()
            y = run_chain([_ident, _ident], x)
            ts.append(time.perf_counter() - t0)
        assert np.array_equal(y, x)
        out["cpu_port"] = {"ops_per_sec": 2 / _median(ts), "ms_per_op": 1e3 * _median(ts) / 2, "cores": 1, "kind": "port",
                           "sample": f"the same two-op chain through oracle/local_runtime.py (one subprocess per op), median of {cpu_reps}"}
        out["speedup_vs_cpu_port"] = out["f32_1MiB"]["ops_per_sec"] / out["cpu_port"]["ops_per_sec"]
    return out


def _ident(a):
    return a


# ------------------------------------------------------------------------------------------------------------------
# C3 — collective over the process group (all ranks call it)
# ------------------------------------------------------------------------------------------------------------------
def c3_nccl(dist, rank: int, world: int, dev: int, mib: int = 4096, reps: int = 4) -> Optional[Dict[str, Any]]:
    """every rank of the job calls this; rank 0 owns the source slot and returns the record"""
    import torch

    from lzy_b200 import _native as N
    from lzy_b200.fanout import broadcast_slot_nccl
    n = mib << 20
    src = None
    dg = None
    if rank == 0:
        t = torch.empty(n, dtype=torch.uint8, device=f"cuda:{dev}")
        N.fill_pattern(dev, t.data_ptr(), n, 7)
        torch.cuda.synchronize(dev)
        src = N.SlotHandle.create_output("c3/src", dev, n)
        src.set_meta(N.DT_U8, (n,))
        src.pack(t.data_ptr(), n)
        dg = src.commit()
        del t
    bc_ms, total_ms, ok = [], [], True
    for r in range(reps + 1):
        dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        slot, ms = broadcast_slot_nccl(dist, 0, src, dev, f"c3/in{rank}/{r}")
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) * 1e3
        both = torch.tensor([ms, dt], device=f"cuda:{dev}")
        dist.all_reduce(both, op=dist.ReduceOp.MAX)  # time to the LAST consumer
        if r:
            bc_ms.append(float(both[0].item()))
            total_ms.append(float(both[1].item()))
        if rank != 0:
            ok = ok and slot.info().state == N.ST_READY
            slot.close()
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{dev}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank != 0:
        return None
    src.close()
    tb, tt = _median(bc_ms), _median(total_ms)
    recv = world - 1
    return {"config": "C3", "method": "one ncclBroadcast into the consumers' slot frames + digest scan on every consumer",
            "payload_mib": mib, "receivers": recv, "broadcast_ms": tb, "time_to_last_verified_ms": tt,
            "per_receiver_GBps": n / tt / 1e6, "aggregate_GBps": recv * n / tt / 1e6,
            "nvlink_frac_per_receiver": (n / tt / 1e6) / 770.0, "verified_on_every_consumer": bool(flag.item()),
            "bound_ms": n / 770e6}


def c3_in_process(gpus: List[int], mib: int = 4096, reps: int = 3) -> List[Dict[str, Any]]:
    """rank 0 alone, one process driving all GPUs: device-chained pull kernel vs concurrent unicasts"""
    import torch

    from lzy_b200 import _native as N
    n = mib << 20
    g0 = gpus[0]
    src_t = torch.empty(n, dtype=torch.uint8, device=f"cuda:{g0}")
    N.fill_pattern(g0, src_t.data_ptr(), n, 7)
    torch.cuda.synchronize(g0)
    src = N.SlotHandle.create_output("c3/src", g0, n)
    src.pack(src_t.data_ptr(), n)
    dg = src.commit()
    del src_t
    consumers = [g for g in gpus if g != g0]
    out = []

    def run_chain():
        dsts = [N.SlotHandle.open_input(f"c3/d{g}", g, n) for g in consumers]
        t0 = time.perf_counter()
        N.broadcast(src, dsts, 0)
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
        for d, s in zip(dsts, streams):
            d.finish(src, int(s.cuda_stream))
        dt = time.perf_counter() - t0
        ok = all(tuple(d.info().digest) == dg for d in dsts)
        for d in dsts:
            d.close()
        return dt, ok

    for name, fn in (("device-chained pull kernel (one launch per GPU, progress counters in peer memory)", run_chain),
                     (f"{len(consumers)} concurrent unicasts (one consumer pull per GPU, all from the same source)", run_unicasts)):
        ts, ok = [], True
        for r in range(reps + 1):
            dt, k = fn()
            ok = ok and k
            if r:
                ts.append(dt)
        t = _median(ts)
        out.append({"config": "C3", "method": name, "payload_mib": mib, "receivers": len(consumers), "time_to_last_verified_ms": t * 1e3,
                    "per_receiver_GBps": n / t / 1e9, "aggregate_GBps": len(consumers) * n / t / 1e9, "verified_on_every_consumer": bool(ok)})
    src.close()
    return out


# ------------------------------------------------------------------------------------------------------------------
# C4
# ------------------------------------------------------------------------------------------------------------------
def c4(dev: int = 0, rows_list=(569, 1_000_000), reps: int = 5, with_cpu: bool = True) -> Dict[str, Any]:
    import numpy as np
    from sklearn.utils import Bunch

    from lzy_b200 import _native as N
    from lzy_b200.api.v1 import B200Runtime, Lzy, materialize, op

    @op
    def train(b: Bunch) -> np.ndarray:  # stand-in for CatBoost fit (catboost is not in the image): a column reduction
        return b.data.sum(axis=0) + float(b.target.sum())

    out: Dict[str, Any] = {"config": "C4", "workload": "Bunch(data N×30 f64, target N i64, 30 names, 4 KiB DESCR) → train stand-in op → ndarray[30]",
                           "runs": []}
    rt = B200Runtime(gpus=[dev], placement="single")
    lzy = Lzy(runtime=rt)
    for rows in rows_list:
        rng = np.random.default_rng(0)
        bunch = Bunch(data=rng.standard_normal((rows, 30)), target=rng.integers(0, 2, rows),
                      feature_names=[f"f{i}" for i in range(30)], DESCR="d" * 4096)
        want = bunch.data.sum(axis=0) + float(bunch.target.sum())
        ts = []
        n_rep = reps if rows < 100000 else max(2, reps // 2)
        N.lib.lzyb_stats_reset()
        for r in range(n_rep + 1):
            t0 = time.perf_counter()
            with lzy.workflow("c4"):
                res = materialize(train(bunch))
            dt = time.perf_counter() - t0
            if r:
                ts.append(dt)
            rt.clear()
        st = N.stats()
        t = _median(ts)
        nbytes = bunch.data.nbytes + bunch.target.nbytes
        rec = {"rows": rows, "array_bytes": nbytes, "wall_ms": t * 1e3, "ops_per_sec": 1 / t, "GBps": nbytes / t / 1e9,
               "correct": bool(np.array_equal(res, want)),
               "bytes_by_path_per_op": {"h2d_into_frames": st["h2d_bytes"] // (n_rep + 1), "d2h_from_frames": st["d2h_bytes"] // (n_rep + 1),
                                        "host_ring_framed": st["bytes_framed_host"] // (n_rep + 1)}}
        if with_cpu and rows <= 1_000_000:
            from oracle.local_runtime import run_chain
            t0 = time.perf_counter()
            ref = run_chain([_train_plain], bunch)
            dtc = time.perf_counter() - t0
            rec["cpu_port"] = {"wall_ms": dtc * 1e3, "ops_per_sec": 1 / dtc, "cores": 1, "kind": "port", "same_result": bool(np.array_equal(ref, want))}
            rec["speedup_vs_cpu_port"] = dtc / t
        out["runs"].append(rec)
    rt.close()
    return out


def _train_plain(b):
    return b.data.sum(axis=0) + float(b.target.sum())


# ------------------------------------------------------------------------------------------------------------------
# C5
# ------------------------------------------------------------------------------------------------------------------
def c5(gpus: Optional[List[int]] = None, mib: int = 256, reps: int = 4) -> Dict[str, Any]:
    """the diamond on the in-process runtime over `gpus`.  Overlap: every branch op records CUDA events around its
    busy work on its GPU's stream; with the transfer log of the runtime (start / end per edge) this gives the time the
    GPUs spent computing while a transfer into a downstream slot was in flight."""
    from typing import Dict as TDict

    import torch

    from lzy_b200.api.v1 import B200Runtime, Lzy, materialize, op
    meta = {f"k{i}": ("v" * 40, i, float(i)) for i in range(700)}  # ≈ 64 KiB pickled
    compute_ms: List[float] = []

    @op
    def source() -> (torch.Tensor, TDict[str, tuple]):
        return torch.ones((mib << 20) // 4, device="cuda"), meta

    @op
    def branch(t: torch.Tensor, m: TDict[str, tuple]) -> (torch.Tensor, TDict[str, tuple]):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t = t.mul_(1.0)
        for _ in range(12):  # busy work of about one transfer time
            t = t * 1.0
        e1.record()
        compute_events.append((e0, e1))
        return t, m

    @op
    def sink(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, d: torch.Tensor, e: torch.Tensor, f: torch.Tensor,
             g: torch.Tensor) -> float:
        dev = a.device
        return float(sum(float(x.to(dev)[0].item()) for x in (a, b, c, d, e, f, g)))

    compute_events: List[Any] = []
    out: Dict[str, Any] = {"config": "C5", "workload": f"source → 7 branches × 2 ops → sink = 16 ops; {mib} MiB fp32 tensor + ~64 KiB dict per edge",
                           "gpus": None}
    ngpu = None
    for prefetch in (True, False):
        kw = {} if gpus is None else {"gpus": list(gpus)}
        rt = B200Runtime(placement="round_robin", prefetch=prefetch, **kw)
        lzy = Lzy(runtime=rt)
        ts, comp, xfer = [], [], []
        res = None
        for r in range(reps + 1):
            compute_events.clear()
            rt._ensure_ready()
            ngpu = len(rt.gpus)
            for gdev in rt.gpus:
                torch.cuda.synchronize(gdev)
            rt.transfer_log(clear=True)
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
            for gdev in rt.gpus:
                torch.cuda.synchronize(gdev)
            if r:
                ts.append(dt)
                comp.append(sum(a.elapsed_time(b) for a, b in compute_events))
                xfer.append(sum(rec.get("ns", 0) for rec in rt.transfer_log(clear=True) if rec.get("bytes", 0) >= (1 << 20)) / 1e6)
            rt.clear()
        tmed = _median(ts)
        out["prefetch_on" if prefetch else "prefetch_off"] = {
            "wall_ms": tmed * 1e3, "ops_per_sec": 16 / tmed, "correct": bool(res == 7.0),
            "branch_compute_ms_sum": _median(comp), "tensor_transfer_ms_sum": _median(xfer)}
        rt.close()
    out["gpus"] = ngpu
    on, off = out["prefetch_on"], out["prefetch_off"]
    # work that would take compute + transfer back to back ran in `wall`: the part of the smaller of the two that was hidden
    serial_on = on["branch_compute_ms_sum"] + on["tensor_transfer_ms_sum"]
    out["overlap_fraction"] = max(0.0, min(1.0, (off["wall_ms"] - on["wall_ms"]) / max(1e-9, off["wall_ms"])))
    out["overlap_note"] = ("overlap_fraction = (wall without prefetch − wall with prefetch) ÷ wall without prefetch; the sums are per-op CUDA-event "
                           f"compute time and per-edge transfer time across all GPUs (serial sum {serial_on:.1f} ms runs in {on['wall_ms']:.1f} ms of wall)")
    return out


# ------------------------------------------------------------------------------------------------------------------
# the process-per-GPU deployment: channel manager + one worker process per GPU, tasks sent as WorkerApi.Execute
# ------------------------------------------------------------------------------------------------------------------
def worker_processes_chain(gpus: List[int], mib: int = 256, n_ops: int = 8, reps: int = 3) -> Dict[str, Any]:
    """The chain of `_worker_processes_chain` twice: with the control calls of the workers (Bind / TransferCompleted towards the
    channel manager, Describe towards the producer's worker) on the library's unix sockets (csrc/host/local_rpc.cpp) — the
    record — and over gRPC as between stock processes — the A/B, and the fallback should the sockets fail on this box."""
    try:
        rec = _worker_processes_chain(gpus, mib, n_ops, reps, local=True)
    except Exception as e:  # noqa: BLE001
        rec = _worker_processes_chain(gpus, mib, n_ops, reps, local=False)
        rec["local_sockets_error"] = f"{type(e).__name__}: {e}"
        return rec
    try:
        grpc_rec = _worker_processes_chain(gpus, mib, n_ops, reps, local=False)
        rec["control_calls_over_grpc"] = {k: grpc_rec[k] for k in ("wall_ms", "ops_per_sec", "edge_GBps", "bit_exact")}
    except Exception as e:  # noqa: BLE001
        rec["control_calls_over_grpc"] = {"error": f"{type(e).__name__}: {e}"}
    return rec


def _worker_processes_chain(gpus: List[int], mib: int, n_ops: int, reps: int, local: bool) -> Dict[str, Any]:
    """An n-op identity chain over `mib`-MiB float32 tensors on the deployment shape of the reference (a worker process
    per GPU — lzy/worker/.../Worker.java:36-56 — and the channel manager as its own process), driven over the kept wire
    contract: LzyChannelManagerPrivate.GetOrCreate for the channels, WorkerApi.Init / Execute + LongRunningService.Get per
    task.  Op k runs on the worker of GPU k mod N; every edge is produced in one process and consumed in another: the
    consumer's transfer factory asks the producer's worker for the frame descriptor (LzybPeerApi.Describe), maps it with
    CUDA IPC and pulls over NVLink with digest verification (TMA ring).  Wall time from the first Execute to the last
    response, the check op compares every element on its GPU."""
    import socket
    import subprocess
    import sys
    from pathlib import Path

    import torch

    from lzy_b200.api.v1.startup import ProcessingRequest
    from lzy_b200.api.v1.utils.pickle import pickle
    from lzy_b200.grpc_api import ChannelManagerPrivateClient, LongRunningClient, WorkerApiClient, messages as M
    root = Path(__file__).resolve().parent
    n_elem = (mib << 20) // 4

    def free_port() -> int:
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            return sk.getsockname()[1]

    def make() -> torch.Tensor:
        import torch as t_
        return t_.arange(n_elem, dtype=t_.float32, device="cuda")

    def hop(t: torch.Tensor) -> torch.Tensor:
        return t

    def check(t: torch.Tensor) -> int:
        import torch as t_
        assert t.is_cuda and t.numel() == n_elem
        assert t_.equal(t, t_.arange(n_elem, dtype=t_.float32, device=t.device))
        return int(t.device.index)

    procs, clients = [], {}
    import shutil
    import tempfile
    sockdir = tempfile.mkdtemp(prefix="lzyb-", dir="/tmp")  # unix socket paths are short
    try:
        cm = subprocess.Popen([sys.executable, "-m", "lzy_b200.grpc_api", "channel-manager", "--port", "0"] +
                              (["--local-socket", f"{sockdir}/cm.sock"] if local else []), cwd=str(root),
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        procs.append(cm)
        cm_target = cm.stdout.readline().strip().rsplit(" ", 1)[-1]
        targets = []
        for g in gpus:
            port = free_port()
            w = subprocess.Popen([sys.executable, "-m", "lzy_b200.grpc_api", "worker", "--gpu", str(g), "--port", str(port),
                                  "--channel-manager", cm_target] +
                                 (["--channel-manager-local", f"{sockdir}/cm.sock", "--local-socket", f"{sockdir}/w{g}.sock"] if local else []),
                                 cwd=str(root), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            procs.append(w)
            line = w.stdout.readline()
            if not line.startswith("worker for GPU"):
                raise RuntimeError(f"worker for GPU {g} did not start: {line!r}")
            targets.append(f"127.0.0.1:{port}")
            clients[g] = (WorkerApiClient(targets[-1]), LongRunningClient(targets[-1]))
            clients[g][0].call("Init", M.InitRequest(user_id="u", workflow_name="wf"), timeout=180.0)
        priv = ChannelManagerPrivateClient(cm_target)

        def run_once(tag: str) -> float:
            def channel(uri):
                sp = M.PeerDescription.StoragePeer(storage_uri=uri)
                return priv.call("GetOrCreate", M.GetOrCreateRequest(execution_id=f"exec-{tag}", workflow_name="wf", user_id="u",
                                                                     consumer=sp)).channel_id

            def execute(g, task_id, fn, args_paths, outs, slots, assign):
                wapi, ops = clients[g]
                request = ProcessingRequest({}, [], fn, args_paths, {}, outs, (tuple, f"/{task_id}/exc"), lazy_arguments=False)
                req = M.ExecuteRequest(task_id=task_id, execution_id=f"exec-{tag}", workflow_name="wf", user_id="u")
                req.task_desc.operation.command = "python -u startup.py " + pickle(request)
                for name, direction in slots:
                    req.task_desc.operation.slots.add(name=name, direction=direction)
                for name, chan in assign.items():
                    req.task_desc.slot_assignments.add(slot_name=name, channel_id=chan)
                op_msg = wapi.execute_and_wait(req, timeout=90.0)  # answered with the completed operation (lzyb-wait-ms)
                if not op_msg.done:
                    # one Get, answered by the worker when the task completes (`lzyb-wait-ms`, grpc_api/server.py): polling
                    # every 0.5 ms took the task's own interpreter ≈ 0.3 ms per poll (tools/worker_chain_cpu.py)
                    op_msg = ops.wait(op_msg.id, timeout=90.0)
                resp = M.ExecuteResponse()
                assert op_msg.WhichOneof("result") == "response" and op_msg.response.Unpack(resp)
                if resp.rc != 0:
                    raise RuntimeError(f"{task_id}: rc {resp.rc}: {resp.description}")

            chans = [channel(f"hbm://{tag}/t{k}/ret") for k in range(n_ops + 1)]
            excs = [channel(f"hbm://{tag}/t{k}/exc") for k in range(n_ops + 2)]
            res = channel(f"hbm://{tag}/check/ret")
            t0 = time.perf_counter()
            execute(gpus[0], f"{tag}-t0", make, [], [(torch.Tensor, "/t0/ret")], [("/t0/ret", 2), (f"/{tag}-t0/exc", 2)],
                    {"/t0/ret": chans[0], f"/{tag}-t0/exc": excs[0]})
            for k in range(1, n_ops + 1):
                g = gpus[k % len(gpus)]
                execute(g, f"{tag}-t{k}", hop, [(torch.Tensor, f"/t{k}/arg")], [(torch.Tensor, f"/t{k}/ret")],
                        [(f"/t{k}/arg", 1), (f"/t{k}/ret", 2), (f"/{tag}-t{k}/exc", 2)],
                        {f"/t{k}/arg": chans[k - 1], f"/t{k}/ret": chans[k], f"/{tag}-t{k}/exc": excs[k]})
            dt = time.perf_counter() - t0
            g = gpus[(n_ops + 1) % len(gpus)]
            execute(g, f"{tag}-check", check, [(torch.Tensor, "/check/arg")], [(int, "/check/ret")],
                    [("/check/arg", 1), ("/check/ret", 2), (f"/{tag}-check/exc", 2)],
                    {"/check/arg": chans[n_ops], "/check/ret": res, f"/{tag}-check/exc": excs[n_ops + 1]})
            priv.call("DestroyAll", M.DestroyAllRequest(execution_id=f"exec-{tag}"))
            return dt

        run_once("warm")
        ts = [run_once(f"r{i}") for i in range(reps)]
        priv.close()
        t = _median(ts)
        return {"config": "C2 on worker processes", "gpus": len(gpus), "processes": len(gpus) + 1, "payload_mib": mib, "ops": n_ops,
                "wall_ms": t * 1e3, "ops_per_sec": n_ops / t, "edge_GBps": n_ops * (mib << 20) / t / 1e9, "bit_exact": True,
                "control_calls": ("unix sockets served by the library's threads (local_rpc.cpp): Bind / TransferCompleted to the channel "
                                  "manager, Describe to the producer's worker" if local else "gRPC, as between stock processes"),
                "wire": "WorkerApi.Init/Execute (answered on completion: lzyb-wait-ms; LongRunningService.Get honours the same header) + LzyChannelManagerPrivate over gRPC; "
                        "payload over NVLink through CUDA IPC (consumer pull, TMA ring, digest verified on landing)"}
    finally:
        for wapi, ops in clients.values():
            try:
                wapi.close()
                ops.close()
            except Exception:
                pass
        for pr in reversed(procs):
            pr.terminate()
        for pr in procs:
            try:
                pr.wait(timeout=20)
            except Exception:
                pr.kill()
        shutil.rmtree(sockdir, ignore_errors=True)
