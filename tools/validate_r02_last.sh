#!/usr/bin/env bash
# the round's last GPU call (1 GPU):  /usr/local/graft/bin/gpurun --timeout 200 -- 'bash tools/validate_r02_last.sh'
# the cross-process tests on ONE GPU (first run there; twice alone, once inside the suite), every -m gpu test as the driver
# runs them, and a short bench line of the final tree
set -u
mkdir -p gpurun_out
for i in 1 2; do
  timeout 60 python -m pytest tests/test_multiprocess_gpu.py -m gpu -q --durations=3 > gpurun_out/pytest_multiprocess_1gpu_$i.log 2>&1; tail -6 gpurun_out/pytest_multiprocess_1gpu_$i.log | cut -c1-200
done
timeout 120 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_last.log 2>&1; tail -2 gpurun_out/pytest_gpu_last.log
timeout 60 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>gpurun_out/bench_n1_last.err | grep '^{' > gpurun_out/bench_n1_last.json
python - <<'PY'
import json
r = json.load(open("gpurun_out/bench_n1_last.json"))
print("bench:", round(r["value"], 1), "| ms/step", round(r["ms_per_step"], 3), "| frac", round(r["roofline"]["frac"], 4), "| step", round(r["roofline"]["step_level"]["frac"], 3),
      "| e2e", round(r["e2e"]["value"], 1), "| api", round(r["api"]["device_resident"]["ms_per_step"], 2), "| parity", r["parity"]["bit_exact"], "| clocks", r["clocks"]["sm_mhz"], r["clocks"]["reasons"])
PY
