"""One worker *process* per GPU, a channel manager process, tasks sent as WorkerApi.Execute — the deployment shape of
the reference (a worker VM per task pool, lzy/worker/.../Worker.java) on one box.  Task `make` runs on the GPU-0 worker
and leaves a tensor in its output slot; task `check` runs on the GPU-1 worker, whose input slot names a peer that
lives in the other process: its transfer factory fetches the frame's descriptor over LzybPeerApi.Describe, maps it
with CUDA IPC (lzyb_peer_import) and pulls over NVLink with digest verification.  Needs ≥ 2 GPUs
(`gpurun --gpus 2`).  NOTE: written at the end of round 1 after the GPU budget was spent — the same flow passes
across three processes on CPU with the in-memory backend (tests/test_grpc_surface.py), this file is its first
run on CUDA IPC."""
import socket
import subprocess
import sys
import time
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs at least 2 GPUs", allow_module_level=True)

from lzy_b200.api.v1.startup import ProcessingRequest  # noqa: E402
from lzy_b200.api.v1.utils.pickle import pickle  # noqa: E402
from lzy_b200.grpc_api import (ChannelManagerPrivateClient, LongRunningClient, WorkerApiClient, messages as M)  # noqa: E402
from lzy_b200.grpc_api.protos import GetOperationRequest  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent
N_ELEM = 8 << 20  # 32 MiB of float32


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def execute(target, task_id, op, args_paths, outs, slots, assign):
    wapi, ops = WorkerApiClient(target), LongRunningClient(target)
    try:
        exc = f"/{task_id}/exc"
        request = ProcessingRequest({}, [], op, args_paths, {}, outs, (tuple, exc), lazy_arguments=False)
        req = M.ExecuteRequest(task_id=task_id, execution_id="exec-mpg", workflow_name="wf", user_id="u")
        req.task_desc.operation.command = "python -u startup.py " + pickle(request)
        for name, direction in slots:
            req.task_desc.operation.slots.add(name=name, direction=direction)
        for name, chan in assign.items():
            req.task_desc.slot_assignments.add(slot_name=name, channel_id=chan)
        op_msg = wapi.call("Execute", req)
        deadline = time.time() + 120
        while not op_msg.done:
            assert time.time() < deadline
            time.sleep(0.01)
            op_msg = ops.call("Get", GetOperationRequest(operation_id=op_msg.id))
        resp = M.ExecuteResponse()
        assert op_msg.WhichOneof("result") == "response" and op_msg.response.Unpack(resp)
        return resp
    finally:
        wapi.close()
        ops.close()


def test_consumer_process_pulls_over_cuda_ipc_from_producer_process():
    n_elem = N_ELEM

    # defined here so that cloudpickle ships them by value (the worker processes do not import this test module)
    def make() -> torch.Tensor:
        import torch as t_
        return t_.arange(n_elem, dtype=t_.float32, device="cuda")

    def check(t: torch.Tensor) -> int:
        import torch as t_
        assert t.is_cuda and t.numel() == n_elem
        assert t_.equal(t, t_.arange(n_elem, dtype=t_.float32, device=t.device))
        return int(t.device.index)

    procs = []
    try:
        cm = subprocess.Popen([sys.executable, "-m", "lzy_b200.grpc_api", "channel-manager", "--port", "0"], cwd=str(ROOT),
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        procs.append(cm)
        cm_target = cm.stdout.readline().strip().rsplit(" ", 1)[-1]
        targets = []
        for gpu in (0, 1):
            port = free_port()
            w = subprocess.Popen([sys.executable, "-m", "lzy_b200.grpc_api", "worker", "--gpu", str(gpu), "--port", str(port),
                                  "--channel-manager", cm_target], cwd=str(ROOT), stdout=subprocess.PIPE,
                                 stderr=subprocess.DEVNULL, text=True)
            procs.append(w)
            line = w.stdout.readline()
            assert line.startswith("worker for GPU"), line
            targets.append(f"127.0.0.1:{port}")
            init = WorkerApiClient(targets[-1])
            init.call("Init", M.InitRequest(user_id="u", workflow_name="wf"), timeout=120.0)
            init.close()

        priv = ChannelManagerPrivateClient(cm_target)

        def channel(uri):
            sp = M.PeerDescription.StoragePeer(storage_uri=uri)
            return priv.call("GetOrCreate", M.GetOrCreateRequest(execution_id="exec-mpg", workflow_name="wf", user_id="u",
                                                                 consumer=sp)).channel_id

        data = channel("hbm://mpg/make/ret")
        r = execute(targets[0], "make", make, [], [(torch.Tensor, "/make/ret")],
                    [("/make/ret", 2), ("/make/exc", 2)], {"/make/ret": data, "/make/exc": channel("hbm://mpg/make/exc")})
        assert (r.rc, r.description) == (0, "Success")
        r = execute(targets[1], "check", check, [(torch.Tensor, "/check/arg")], [(int, "/check/ret")],
                    [("/check/arg", 1), ("/check/ret", 2), ("/check/exc", 2)],
                    {"/check/arg": data, "/check/ret": channel("hbm://mpg/check/ret"), "/check/exc": channel("hbm://mpg/check/exc")})
        assert (r.rc, r.description) == (0, "Success")  # the op body compared every element on GPU 1
        priv.close()
    finally:
        for p in reversed(procs):
            p.terminate()
        for p in procs:
            try:
                p.wait(timeout=20)
            except subprocess.TimeoutExpired:
                p.kill()
