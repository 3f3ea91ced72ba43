"""WorkerApi.Execute with the real mount: arguments are views of the HBM frames `beforeExecution` delivered,
results are packed into the output slots' frames by the CUDA pack kernel, `afterExecution` hands them to the
HBM frame store (the storage peer of this box).  Same staging as tests/test_worker.py, on cuda:0."""
import uuid

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from lzy_b200 import _host as H  # noqa: E402
from lzy_b200.api.v1.b200.store import FrameStore  # noqa: E402
from lzy_b200.api.v1.startup import ProcessingRequest  # noqa: E402
from lzy_b200.api.v1.utils.pickle import pickle  # noqa: E402
from lzy_b200.serialization.registry import LzySerializerRegistry  # noqa: E402
from lzy_b200.worker import DIR_INPUT, DIR_OUTPUT, SlotSpec, TaskSpec, Worker  # noqa: E402


def scale_and_describe(t: torch.Tensor, k: int) -> (torch.Tensor, dict):
    assert t.is_cuda
    return t * k, {"sum": float(t.sum().item()), "shape": tuple(t.shape)}


def explode(t: torch.Tensor) -> torch.Tensor:
    raise RuntimeError("kaboom on " + str(t.device))


@pytest.fixture
def stack(native):
    cm = H.ChannelManager(seed=9)
    H.pump_start(cm)
    worker = Worker(cm, 0, "lzyb://gpu-worker-0")
    worker.init("uid", "wf")
    store, ser = FrameStore(0), LzySerializerRegistry()
    prefix = f"hbm://wtest-{uuid.uuid4().hex[:8]}/"
    yield cm, worker, store, ser, prefix
    worker.close()
    store.clear(prefix)
    H.pump_stop(cm)
    cm.close()


def build(cm, store, ser, prefix, op, args, out_types, task_id):
    slots, assign, arg_paths, outs = [], {}, [], []
    for i, (typ, value) in enumerate(args):
        uri = prefix + store.put(prefix, value, typ, ser, name=f"arg{i}")
        path = f"/{task_id}/a{i}"
        assign[path] = cm.get_or_create("uid", "wf", "exec-g", producer_uri=uri)
        slots.append(SlotSpec(path, DIR_INPUT))
        arg_paths.append((typ, path))
    for i, typ in enumerate(out_types):
        path = f"/{task_id}/r{i}"
        assign[path] = cm.get_or_create("uid", "wf", "exec-g", consumer_uri=f"{prefix}{task_id}/r{i}")
        slots.append(SlotSpec(path, DIR_OUTPUT))
        outs.append((typ, path))
    exc = f"/{task_id}/exc"
    assign[exc] = cm.get_or_create("uid", "wf", "exec-g", consumer_uri=f"{prefix}{task_id}/exc")
    slots.append(SlotSpec(exc, DIR_OUTPUT))
    req = ProcessingRequest({}, ser.imports(), op, arg_paths, {}, outs, (tuple, exc), lazy_arguments=False)
    return TaskSpec(task_id, "exec-g", "wf", "uid", "python -u startup.py " + pickle(req), slots, assign)


def test_execute_on_the_gpu_worker(stack, native):
    cm, worker, store, ser, prefix = stack
    t = torch.arange(3 << 18, dtype=torch.float32, device="cuda:0").reshape(3, -1)
    before = native.stats()
    task = build(cm, store, ser, prefix, scale_and_describe, [(torch.Tensor, t), (int, 3)], (torch.Tensor, dict), "t1")
    done = worker.wait_operation(worker.execute(task).id)
    assert done.error is None and (done.response.rc, done.response.description) == (0, "Success")
    got = store.get(f"{prefix}t1/r0", torch.Tensor, ser)
    assert got.is_cuda and torch.equal(got, t * 3)
    meta = store.get(f"{prefix}t1/r1", dict, ser)
    assert meta == {"sum": float(t.sum().item()), "shape": (3, 1 << 18)}
    assert store.get(f"{prefix}t1/exc", type(None), ser) is None
    after = native.stats()
    assert after["pack_launches"] > before["pack_launches"] and after["pull_launches"] > before["pull_launches"]
    # chained: t1's result channel feeds t2 on the same worker, slot to slot
    t2 = build(cm, store, ser, prefix, scale_and_describe, [(torch.Tensor, t), (int, 2)], (torch.Tensor, dict), "t2")
    t2.slot_assignments["/t2/a0"] = task.slot_assignments["/t1/r0"]
    done2 = worker.wait_operation(worker.execute(t2).id)
    assert done2.response.rc == 0
    assert torch.equal(store.get(f"{prefix}t2/r0", torch.Tensor, ser), t * 6)
    worker.release_execution("exec-g")


def test_failing_op_on_the_gpu_worker(stack):
    cm, worker, store, ser, prefix = stack
    t = torch.ones(1024, device="cuda:0")
    task = build(cm, store, ser, prefix, explode, [(torch.Tensor, t)], (torch.Tensor,), "t3")
    done = worker.wait_operation(worker.execute(task).id)
    assert done.response.rc == 1
    import time
    deadline = time.time() + 5
    while not store.exists(f"{prefix}t3/exc") and time.time() < deadline:
        time.sleep(0.01)
    exc = store.get(f"{prefix}t3/exc", tuple, ser)
    assert exc[0] is RuntimeError and "kaboom on cuda:0" in str(exc[1])
    worker.release_execution("exec-g")
