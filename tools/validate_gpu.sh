#!/usr/bin/env bash
# One gpurun call's worth of validation on a 1-GPU box (≈ 2.5 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/validate_gpu.sh'
# -m gpu tests (the multi-GPU file skips itself), smoke(), the default bench line and a short reference arm.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
# tests written after the previous round's GPU minutes were spent: first run, not -x, full report
LZYB_UNVALIDATED=1 timeout 300 python -m pytest tests/test_workflow_gpu.py -m gpu -q -k 'nested_workflow_on_the_same_runtime or identical_cached' 2>&1 | tail -15
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-100
timeout 400 python bench.py 2>gpurun_out/bench.err | grep '^{' > gpurun_out/bench.json
python - <<'PY'
import json
r = json.load(open("gpurun_out/bench.json"))
print("bench:", round(r["value"], 1), r["unit"], "| ms/step", round(r["ms_per_step"], 3), "| roofline frac", round(r["roofline"]["frac"], 4),
      "| e2e", round(r["e2e"]["value"], 1), "| e2e duplex", (r["e2e"].get("duplex") or {}).get("value"),
      "| launches", r["gpu_launches"], "| clocks", r["clocks"]["sm_mhz"], r["clocks"]["reasons"])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-200
# per-op phase costs through the API (host changes of the previous round's tail show up here)
LZYB_TRACE=1 timeout 200 python tools/profile_chain.py > gpurun_out/profile_chain.txt 2>&1; tail -12 gpurun_out/profile_chain.txt
# A/B of the recycled slot threads (LZYB_SLOT_THREADS=fresh = one new thread per slot, as before)
LZYB_SLOT_THREADS=fresh LZYB_TRACE=1 timeout 200 python tools/profile_chain.py > gpurun_out/profile_chain_fresh_threads.txt 2>&1; tail -4 gpurun_out/profile_chain_fresh_threads.txt
