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
    base = dict(gpus=1, steps=1, warmup=0, impl="reference", payload_mib=1024, engine="auto", dma_last_mib=0, dma_growth=6,
                ctas_per_sm=0, no_cpu_baseline=False, no_e2e=False)
    base.update(kw)
    return argparse.Namespace(**base)


def test_reference_arm_prints_the_agreed_line(bench, monkeypatch, capsys):
    monkeypatch.setattr(bench, "pick_cpu_sample", lambda total_steps, budget_s=150.0: 1 << 20)
    monkeypatch.setenv("RANK", "0")
    bench.run_reference(args())
    (line,) = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    r = json.loads(line)
    assert r["impl"] == "reference" and r["metric"] == "slot_to_slot_GBps" and r["unit"] == "GB/s"
    assert r["higher_is_better"] is True and r["scaling"] == "weak" and r["vs_baseline"] is None
    assert (r["n_gpus"], r["steps"], r["warmup"]) == (1, 1, 0) and r["value"] > 0 and r["ms_per_step"] > 0
    assert r["dtype"] == "u8" and r["data"] == "synthetic" and "workload" in r["config"] and r["gpu_launches"] == 0
    cb = r["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == r["value"] and "subprocess per op" in cb["sample"]
    assert r["e2e"] == {"value": r["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # value is payload bytes delivered per second over the 8 hops of the chain
    assert r["value"] == pytest.approx(8 * (1 << 20) / (r["ms_per_step"] * 1e-3) / 1e9, rel=1e-6)


def test_reference_arm_runs_on_rank_0_only(bench, monkeypatch, capsys):
    monkeypatch.setattr(bench, "pick_cpu_sample", lambda *a, **k: pytest.fail("rank 1 must not do any work"))
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(args(gpus=2))
    assert capsys.readouterr().out == ""


def test_sample_picker_stays_within_its_budget(bench, monkeypatch):
    costs = {64 << 10: 0.8, 32 << 20: 1.6}  # a host where a process spawn costs 0.1 s and 32 MiB add 0.1 s per op
    monkeypatch.setattr(bench, "cpu_chain_seconds", lambda nbytes, n_ops=8: costs[nbytes])
    assert bench.pick_cpu_sample(4, budget_s=150.0) == 512 << 20   # 37.5 s per step: the cap
    assert bench.pick_cpu_sample(33, budget_s=150.0) == 32 << 20   # 4.5 s per step
    assert bench.pick_cpu_sample(1000, budget_s=150.0) == 1 << 20  # never below 1 MiB


def test_product_arm_refuses_to_run_without_a_gpu(bench, monkeypatch):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for GPU-less boxes")
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    with pytest.raises(SystemExit, match="needs a CUDA device"):
        bench.run_ours(args(impl="ours"))
