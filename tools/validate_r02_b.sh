#!/usr/bin/env bash
# second 1-GPU call of the round's tail: the N = 1 chain with commits in stream order (and the blocking A/B), C1 under cProfile
set -u
mkdir -p gpurun_out
timeout 420 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/bench_n1_so.err | grep '^{' > gpurun_out/bench_n1_so.json
python - <<'PY'
import json
r = json.load(open("gpurun_out/bench_n1_so.json"))
print("bench:", round(r["value"], 1), r["unit"], "| ms/step", round(r["ms_per_step"], 3), "| frac", round(r["roofline"]["frac"], 4),
      "| step frac", round(r["roofline"]["step_level"]["frac"], 4), "| share", round(r["roofline"]["kernel_share_of_step"], 3),
      "| e2e", round(r["e2e"]["value"], 1), round(r["e2e"]["ms_per_step"], 2), "ms | cabi", round(r["e2e"]["c_abi_only"]["ms_per_step"], 2),
      "| api dev", round(r["api"]["device_resident"]["ms_per_step"], 2),
      "| launches", r["gpu_launches"], "| clocks", r["clocks"]["sm_mhz"], r["clocks"]["reasons"], "| parity", r["parity"]["bit_exact"], r["parity"]["corruption_detected"])
for c in r["extra"]["configs"]:
    print("  extra:", json.dumps(c)[:200])
PY
tail -3 gpurun_out/bench_n1_so.err
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --blocking-commit --no-e2e --no-extras --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_n1_blocking.json
python -c "
import json; r = json.load(open('gpurun_out/bench_n1_blocking.json')); print('blocking:', round(r['value'], 1), round(r['ms_per_step'], 3))"
timeout 200 python tools/profile_api.py tottime gpurun_out/c1.prof > gpurun_out/profile_api_c1_tottime.txt 2>&1; head -3 gpurun_out/profile_api_c1_tottime.txt
timeout 100 python -m pytest tests/test_kernels_gpu.py tests/test_workflow_gpu.py -m gpu -q 2>&1 | tail -2
