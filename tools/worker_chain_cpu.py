#!/usr/bin/env python
"""Per-op control-plane cost of the process-per-GPU deployment, measured WITHOUT a GPU: the channel manager and two worker
processes (tests/support/remote_worker_api.py — the product's worker over the in-memory test backend) driven exactly as
bench_configs.worker_processes_chain drives the GPU workers (GetOrCreate, WorkerApi.Execute, LongRunningService.Get), an n-op
identity chain over a small `bytes` payload.  What remains when the payload is negligible is the wire + staging cost per
op — the 5–10 ms that the GPU leg reports next to a 0.4 ms NVLink transfer.  A tuning aid; prints one JSON line."""
import json
import os
import socket
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from lzy_b200.api.v1.startup import ProcessingRequest  # noqa: E402
from lzy_b200.api.v1.utils.pickle import pickle  # noqa: E402
from lzy_b200.grpc_api import ChannelManagerPrivateClient, LongRunningClient, WorkerApiClient, messages as M  # noqa: E402
from lzy_b200.grpc_api.protos import GetOperationRequest  # noqa: E402

N_BYTES = 64 << 10
POLL_S = float(os.environ["POLL_S"]) if os.environ.get("POLL_S") else None  # POLL_S=0.0005: poll Get like a stock client
LOCAL = os.environ.get("LOCAL", "1") != "0"  # LOCAL=0: matcher calls and Describe over gRPC, as between stock processes


def make() -> bytes:
    return bytes(range(256)) * (N_BYTES // 256)


def hop(b: bytes) -> bytes:
    return b


def check(b: bytes) -> int:
    assert b == bytes(range(256)) * (N_BYTES // 256)
    return len(b)


def free_port() -> int:
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def main(n_ops: int = 8, reps: int = 5, n_workers: int = 2) -> None:
    procs, clients = [], []
    tmp = tempfile.mkdtemp(prefix="lzyb-wc-", dir="/dev/shm")
    try:
        cm_sock = f"{tmp}/cm.sock"
        cm = subprocess.Popen([sys.executable, "-m", "lzy_b200.grpc_api", "channel-manager", "--port", "0"] +
                              (["--local-socket", cm_sock] if LOCAL else []), cwd=str(ROOT),
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        procs.append(cm)
        cm_target = cm.stdout.readline().strip().rsplit(" ", 1)[-1]
        for _ in range(n_workers):
            port = free_port()
            w = subprocess.Popen([sys.executable, str(ROOT / "tests/support/remote_worker_api.py"), cm_target, str(port)] +
                                 ([cm_sock, f"{tmp}/w{len(clients)}.sock"] if LOCAL else []),
                                 cwd=str(ROOT), stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                 stderr=open(f"/tmp/lzyb_worker_{len(clients)}.err", "w"), text=True)
            procs.append(w)
            line = w.stdout.readline()
            assert line.startswith("READY"), line
            t = f"127.0.0.1:{port}"
            clients.append((WorkerApiClient(t), LongRunningClient(t)))
            clients[-1][0].call("Init", M.InitRequest(user_id="u", workflow_name="wf"))
        priv = ChannelManagerPrivateClient(cm_target)
        polls = [0]

        def run_once(tag: str):
            def channel(name):
                sp = M.PeerDescription.StoragePeer(storage_uri=f"file://{tmp}/{tag}.{name}")
                return priv.call("GetOrCreate", M.GetOrCreateRequest(execution_id=f"exec-{tag}", workflow_name="wf", user_id="u",
                                                                     consumer=sp)).channel_id

            def execute(i, task_id, fn, args_paths, outs, slots, assign):
                wapi, ops = clients[i % n_workers]
                request = ProcessingRequest({}, [], fn, args_paths, {}, outs, (tuple, f"/{task_id}/exc"), lazy_arguments=False)
                req = M.ExecuteRequest(task_id=task_id, execution_id=f"exec-{tag}", workflow_name="wf", user_id="u")
                req.task_desc.operation.command = "python -u startup.py " + pickle(request)
                for name, direction in slots:
                    req.task_desc.operation.slots.add(name=name, direction=direction)
                for name, chan in assign.items():
                    req.task_desc.slot_assignments.add(slot_name=name, channel_id=chan)
                if POLL_S is None:  # the product's way: Execute answered with the completed operation (else one waiting Get)
                    op_msg = wapi.execute_and_wait(req, timeout=120.0)
                    if not op_msg.done:
                        polls[0] += 1
                        op_msg = ops.wait(op_msg.id)
                else:
                    op_msg = wapi.call("Execute", req)
                while not op_msg.done:  # a stock client's way
                    time.sleep(POLL_S)
                    polls[0] += 1
                    op_msg = ops.call("Get", GetOperationRequest(operation_id=op_msg.id))
                resp = M.ExecuteResponse()
                assert op_msg.WhichOneof("result") == "response" and op_msg.response.Unpack(resp)
                assert resp.rc == 0, resp.description

            t0 = time.perf_counter()
            chans = [channel(f"t{k}.ret") for k in range(n_ops + 1)]
            excs = [channel(f"t{k}.exc") for k in range(n_ops + 2)]
            res = channel("check.ret")
            t1 = time.perf_counter()
            execute(0, f"{tag}-t0", make, [], [(bytes, "/t0/ret")], [("/t0/ret", 2), (f"/{tag}-t0/exc", 2)],
                    {"/t0/ret": chans[0], f"/{tag}-t0/exc": excs[0]})
            t2 = time.perf_counter()
            per = []
            for k in range(1, n_ops + 1):
                s = time.perf_counter()
                execute(k, f"{tag}-t{k}", hop, [(bytes, f"/t{k}/arg")], [(bytes, f"/t{k}/ret")],
                        [(f"/t{k}/arg", 1), (f"/t{k}/ret", 2), (f"/{tag}-t{k}/exc", 2)],
                        {f"/t{k}/arg": chans[k - 1], f"/t{k}/ret": chans[k], f"/{tag}-t{k}/exc": excs[k]})
                per.append(time.perf_counter() - s)
            t3 = time.perf_counter()
            execute(n_ops + 1, f"{tag}-check", check, [(bytes, "/check/arg")], [(int, "/check/ret")],
                    [("/check/arg", 1), ("/check/ret", 2), (f"/{tag}-check/exc", 2)],
                    {"/check/arg": chans[n_ops], "/check/ret": res, f"/{tag}-check/exc": excs[n_ops + 1]})
            priv.call("DestroyAll", M.DestroyAllRequest(execution_id=f"exec-{tag}"))
            return {"channels_ms": 1e3 * (t1 - t0), "first_ms": 1e3 * (t2 - t1), "chain_ms": 1e3 * (t3 - t2),
                    "per_op_ms": [round(1e3 * x, 3) for x in per]}

        run_once("warm")
        runs = []
        for i in range(reps):
            polls[0] = 0
            r = run_once(f"r{i}")
            r["polls"] = polls[0]
            runs.append(r)
        runs.sort(key=lambda r: r["chain_ms"])
        med = runs[len(runs) // 2]
        print(json.dumps({"what": "worker-process chain on CPU (in-memory backend)", "ops": n_ops, "workers": n_workers,
                          "payload_bytes": N_BYTES,
                          "control_calls": "unix sockets of the library (local_rpc.cpp)" if LOCAL else "gRPC", "median": med, "ms_per_op": med["chain_ms"] / n_ops}))
        priv.close()
    finally:
        for wapi, ops in clients:
            try:
                wapi.close(), ops.close()
            except Exception:
                pass
        for p in procs[1:]:
            try:
                p.stdin.write("bye\n")
                p.stdin.flush()
            except Exception:
                pass
        for p in procs[1:]:
            try:
                p.wait(timeout=20)
            except subprocess.TimeoutExpired:
                p.kill()
        procs[0].terminate()
        procs[0].wait(timeout=15)
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
