#!/usr/bin/env bash
# The round's last 1-GPU validation in one gpurun call, most valuable record first:
#   /usr/local/graft/bin/gpurun --timeout 840 -- 'bash tools/validate_r02_final.sh'
# default bench line (driver's arguments) → smoke() → C1 native/Python split → every -m gpu test → ncu launch list.
set -u
mkdir -p gpurun_out
timeout 420 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/bench_n1_final.err | grep '^{' > gpurun_out/bench_n1_final.json
python - <<'PY'
import json
r = json.load(open("gpurun_out/bench_n1_final.json"))
print("bench:", round(r["value"], 1), r["unit"], "| ms/step", round(r["ms_per_step"], 3), "| frac", round(r["roofline"]["frac"], 4),
      "| e2e", round(r["e2e"]["value"], 1), round(r["e2e"]["ms_per_step"], 2), "ms | api dev", round(r["api"]["device_resident"]["ms_per_step"], 2),
      "| launches", r["gpu_launches"], "| clocks", r["clocks"]["sm_mhz"], r["clocks"]["reasons"], "| parity", r["parity"]["bit_exact"])
for c in r["extra"]["configs"]:
    print("  extra:", json.dumps(c)[:260])
PY
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-160
timeout 120 python tools/c1_breakdown.py > gpurun_out/c1_breakdown.txt 2>&1; head -14 gpurun_out/c1_breakdown.txt
timeout 600 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu_final.log 2>&1; tail -14 gpurun_out/pytest_gpu_final.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_n1_final.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log | cut -c1-200
