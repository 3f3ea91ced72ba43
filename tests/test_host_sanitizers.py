"""The C++ host logic (channel manager, slot state machines, action pump, storage and spill tiers) under
ThreadSanitizer and under AddressSanitizer + UBSan: tests/native/host_stress.cpp is a native multi-threaded driver
over the in-memory backend, linked against the host sources directly (the CUDA half is stubbed out — it is never
reached).  The reference has no race detector; every slot here is a thread, so this one matters."""
import json
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "tests" / "support" / "_build"
SOURCES = [ROOT / "tests/native/host_stress.cpp", ROOT / "tests/native/dataplane_stubs.cpp",
           *sorted((ROOT / "lzy_b200/csrc/host").glob("*.cpp"))]
HEADERS = [*sorted((ROOT / "include").glob("*.h")), *sorted((ROOT / "lzy_b200/csrc/host").glob("*.h"))]
BACKEND = ROOT / "tests/support/hostmem_backend.c"


def build(name: str, flags):
    exe = BUILD / f"host_stress_{name}"
    deps = SOURCES + HEADERS + [BACKEND, Path(__file__)]
    if exe.exists() and exe.stat().st_mtime >= max(d.stat().st_mtime for d in deps):
        return exe
    BUILD.mkdir(exist_ok=True)
    obj = BUILD / f"hostmem_{name}.o"
    subprocess.run(["gcc", "-O1", "-g", *flags, "-c", f"-I{ROOT / 'include'}", str(BACKEND), "-o", str(obj)], check=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", *flags, f"-I{ROOT / 'include'}", "-o", str(exe),
                    *map(str, SOURCES), str(obj), "-lpthread"], check=True)
    return exe


@pytest.mark.parametrize("name, flags, threads, jobs, over_socket", [
    ("tsan", ["-fsanitize=thread"], 6, 30, False),
    ("asan", ["-fsanitize=address,undefined", "-fno-omit-frame-pointer"], 8, 40, False),
    # the same workout with every matcher call of the slot threads crossing the box-local transport (local_rpc.cpp)
    ("tsan", ["-fsanitize=thread"], 6, 20, True),
    ("asan", ["-fsanitize=address,undefined", "-fno-omit-frame-pointer"], 8, 30, True),
], ids=["tsan", "asan", "tsan-socket", "asan-socket"])
def test_host_logic_under_sanitizers(name, flags, threads, jobs, over_socket, tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = build(name, flags)
    import tempfile
    sockdir = tempfile.mkdtemp(prefix="lzyb-", dir="/tmp") if over_socket else None  # short path: sun_path is 108 bytes
    try:
        r = subprocess.run([str(exe), str(threads), str(jobs), str(tmp_path / "spill")] + ([f"{sockdir}/cm.sock"] if over_socket else []),
                           capture_output=True, text=True, timeout=600)
    finally:
        if sockdir:
            shutil.rmtree(sockdir, ignore_errors=True)
    report = r.stdout + r.stderr
    if "FATAL: ThreadSanitizer" in report or "FATAL: AddressSanitizer" in report or "Shadow memory range interleaves" in report:
        pytest.skip("the sanitizer runtime cannot map its shadow memory in this sandbox: " + report.strip().splitlines()[0])
    assert "ThreadSanitizer" not in report and "AddressSanitizer" not in report and "runtime error" not in report, report[-4000:]
    assert r.returncode == 0, report[-4000:]
    summary = json.loads(r.stdout.strip().splitlines()[-1])
    assert summary == {"threads": threads, "jobs_per_thread": jobs, "failures": 0, "live_slots": 0}
