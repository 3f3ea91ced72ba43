"""bench.py's reporting contract where it can be checked without a GPU: the reference arm prints one JSON line with
the agreed keys (its workload shrunk here to keep the suite fast), ranks other than 0 stay silent, and the product
arm refuses to run without a CUDA device instead of falling back to anything."""
import argparse
import importlib.util
import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def args(**kw):
    base = dict(gpus=1, steps=1, warmup=0, impl="reference", config="c2", payload_mib=1, engine="auto", no_move=False,
                ctas_per_sm=0, no_cpu_baseline=False, no_e2e=False, no_extras=False)
    base.update(kw)
    return argparse.Namespace(**base)


def test_reference_arm_prints_the_agreed_line(bench, monkeypatch, capsys):
    monkeypatch.setenv("RANK", "0")
    bench.run_reference(args())  # the workload shrunk to 1 MiB per edge; the arm itself is unchanged
    (line,) = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    r = json.loads(line)
    assert r["impl"] == "reference" and r["metric"] == "slot_to_slot_GBps" and r["unit"] == "GB/s"
    assert r["higher_is_better"] is True and r["scaling"] == "weak" and r["vs_baseline"] is None
    assert (r["n_gpus"], r["steps"], r["warmup"]) == (1, 1, 0) and r["value"] > 0 and r["ms_per_step"] > 0
    assert r["dtype"] == "u8" and r["data"] == "synthetic" and r["gpu_launches"] == 0
    cb = r["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == r["value"] and "subprocess" in cb["sample"]
    assert r["e2e"] == {"value": r["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a step is a bounded sample of the chain: `sample_ops_per_step` consecutive ops at the full payload of the config
    n_ops = cb["sample_ops_per_step"]
    assert 1 <= n_ops <= 8
    assert r["value"] == pytest.approx(n_ops * (1 << 20) / (r["ms_per_step"] * 1e-3) / 1e9, rel=1e-6)


def test_both_arms_name_the_same_config(bench):
    """`same_config` is decided on this dict: the reference arm times the workload our arm times (1 GiB edges by default),
    and nothing in it says how an arm carries the bytes"""
    ours = bench.workload_config(1 << 30, 1)
    assert ours == bench.workload_config(1 << 30, 1)
    assert ours["workload"] == "linear 8-@op chain, 1 GiB fp32 tensor per edge" and ours["payload_bytes"] == 1 << 30 and ours["ops"] == 8
    assert "carrier" not in ours and "engine" not in ours and "l2" in ours
    assert bench.workload_config(1 << 30, 8)["gpus"] == 8


def test_reference_arm_runs_on_rank_0_only(bench, monkeypatch, capsys):
    monkeypatch.setattr(bench, "cpu_chain_seconds", lambda *a, **k: pytest.fail("rank 1 must not do any work"))
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(args(gpus=2))
    assert capsys.readouterr().out == ""


def test_product_arm_refuses_to_run_without_a_gpu(bench, monkeypatch):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for GPU-less boxes")
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    with pytest.raises(SystemExit, match="needs a CUDA device"):
        bench.run_ours(args(impl="ours"))


def test_rank_0_waits_for_the_other_ranks_processes_not_for_a_process_count(bench):
    """N > 1: before the one-process legs rank 0 waits until the other ranks are gone — a zombie (exited, not yet reaped by
    torchrun) counts as gone, a live process does not, and the wait is bounded"""
    import subprocess
    import time
    quick = subprocess.Popen([sys.executable, "-c", "pass"])
    slow = subprocess.Popen([sys.executable, "-c", "import time; time.sleep(30)"])
    try:
        deadline = time.time() + 10
        while quick.poll() is None and time.time() < deadline:  # reaped here; an unreaped child is covered below
            time.sleep(0.02)
        assert bench.wait_until_alone([quick.pid], timeout_s=5.0) < 1.0
        t0 = time.time()
        waited = bench.wait_until_alone([quick.pid, slow.pid], timeout_s=0.5)
        assert 0.5 <= waited < 2.0 and time.time() - t0 < 2.0
        zombie = subprocess.Popen([sys.executable, "-c", "pass"])  # never polled: stays a zombie until wait()
        assert bench.wait_until_alone([zombie.pid], timeout_s=10.0) < 5.0
        zombie.wait()
    finally:
        slow.kill()
        slow.wait()


def test_worker_process_leg_records_both_transports_and_falls_back(monkeypatch):
    """bench_configs.worker_processes_chain: the record is the run with the control calls on the library's unix sockets, the
    gRPC run rides along as the A/B; sockets that fail on a box leave the gRPC record with the reason next to it"""
    sys.path.insert(0, str(ROOT))
    import bench_configs
    calls = []

    def fake(gpus, mib, n_ops, reps, local):
        calls.append(local)
        return {"config": "C2 on worker processes", "wall_ms": 20.0 if local else 36.0, "ops_per_sec": 1.0, "edge_GBps": 1.0,
                "bit_exact": True, "control_calls": "unix" if local else "gRPC"}
    monkeypatch.setattr(bench_configs, "_worker_processes_chain", fake)
    rec = bench_configs.worker_processes_chain([0, 1])
    assert calls == [True, False] and rec["wall_ms"] == 20.0 and rec["control_calls_over_grpc"]["wall_ms"] == 36.0

    def broken(gpus, mib, n_ops, reps, local):
        if local:
            raise RuntimeError("worker for GPU 0 did not start")
        return fake(gpus, mib, n_ops, reps, local)
    monkeypatch.setattr(bench_configs, "_worker_processes_chain", broken)
    rec = bench_configs.worker_processes_chain([0, 1])
    assert rec["control_calls"] == "gRPC" and "did not start" in rec["local_sockets_error"]

    def grpc_broken(gpus, mib, n_ops, reps, local):
        if not local:
            raise TimeoutError("t3")
        return fake(gpus, mib, n_ops, reps, local)
    monkeypatch.setattr(bench_configs, "_worker_processes_chain", grpc_broken)
    rec = bench_configs.worker_processes_chain([0, 1])
    assert rec["wall_ms"] == 20.0 and "TimeoutError" in rec["control_calls_over_grpc"]["error"]
