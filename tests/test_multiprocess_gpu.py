"""One worker *process* per GPU, a channel manager process, tasks sent as WorkerApi.Execute — the deployment shape of
the reference (a worker VM per task pool, lzy/worker/.../Worker.java) on one box.  Task `make` runs on the GPU-0 worker
and leaves a tensor in its output slot; task `check` runs on another worker, whose input slot names a peer that
lives in the other process: its transfer factory fetches the frame's descriptor over LzybPeerApi.Describe, maps it
with CUDA IPC (lzyb_peer_import) and pulls with digest verification; and the producer-issued push between two processes
(arrival word, one-thread waiter).  With ≥ 2 GPUs (`gpurun --gpus 2`) the second process owns GPU 1 and the bytes cross
NVLink; on a 1-GPU box both processes share GPU 0 — the same descriptors, CUDA-IPC mappings, kernels, arrival word and
verdict path between two processes, without the wire — so that the box the driver tests on sees the cross-process
protocol too.  The same flow passes across three processes on CPU with the in-memory backend (tests/test_grpc_surface.py)."""
import socket
import subprocess
import sys
import time
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)
PEER_GPU = 1 if torch.cuda.device_count() >= 2 else 0  # the GPU of the second process

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
        for gpu in (0, PEER_GPU):
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
        assert (r.rc, r.description) == (0, "Success")  # the op body compared every element on the consumer's GPU
        priv.close()
    finally:
        for p in reversed(procs):
            p.terminate()
        for p in procs:
            try:
                p.wait(timeout=20)
            except subprocess.TimeoutExpired:
                p.kill()


@pytest.mark.parametrize("nbytes", [(96 << 20) + 4096 + 7, (2 << 20) + 1])
def test_producer_process_pushes_into_the_consumer_process_frame(nbytes):
    """The hand-written carrier between worker PROCESSES: the consumer exports its empty input frame (CUDA IPC), the
    producer's process launches the TMA-ring kernel with that frame as destination (digest folded on the way, header,
    verdict and — last — the arrival word written into the frame), the consumer's GPU runs the one-thread
    arrival_wait_kernel.  No copy engine, no cross-process event; the delivered bytes equal the counter-based pattern of
    the oracle, twice (the second delivery reuses the pooled frame under a fresh arrival word)."""
    import numpy as np

    import oracle
    from lzy_b200 import _native as N
    from lzy_b200.devmem import view_bytes

    seed = 77
    child = subprocess.Popen([sys.executable, str(ROOT / "tests" / "support" / "ipc_push_child.py"), "0", str(nbytes), str(seed)],
                             stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, cwd=str(ROOT))
    try:
        producer = N.IpcDesc.from_buffer_copy(bytes.fromhex(child.stdout.readline().strip()))
        assert int(producer.nbytes) == nbytes and int(producer.gpu) == 0
        N.init(0, 0, 64 << 20)  # idempotent: the session may have initialised the library already (all visible GPUs)
        expect = oracle.pattern(nbytes, seed)
        assert tuple(producer.digest) == oracle.digest(expect)
        dma_before = N.stats()["dma_copies"]
        for life in range(2):
            inp = N.SlotHandle.open_input(f"ipcpush/in{life}", PEER_GPU, nbytes)
            sink = inp.sink_export()
            stream = torch.cuda.Stream(device=PEER_GPU)
            sp = int(stream.cuda_stream)
            assert inp.expect_push(producer, 0, 0, sp) == nbytes   # the waiter is spinning on the GPU before a byte has moved
            child.stdin.write(bytes(sink).hex() + "\n")
            child.stdin.flush()
            word, pushes, dma = child.stdout.readline().split()
            assert word == "pushed" and int(pushes) == life + 1 and int(dma) == 0
            inp.finish_expect(producer.digest, sp)
            info = inp.info()
            assert info.state == N.ST_READY and tuple(info.digest) == tuple(producer.digest)
            assert info.dtype == N.DT_U8 and info.ndim == 1 and info.shape[0] == nbytes
            got = view_bytes(inp.device_ptr(), nbytes, PEER_GPU).cpu().numpy()
            assert np.array_equal(got, expect)
            assert N.stats()["dma_copies"] == dma_before
            inp.close()
        # a push that never comes is a peer fault after the scan's timeout, not a hang: covered by the fault tests of
        # the in-process pair (tests/test_multigpu_gpu.py); here only the clean shutdown
        child.stdin.write("quit\n")
        child.stdin.flush()
        assert child.wait(timeout=60) == 0
    finally:
        if child.poll() is None:
            child.kill()
