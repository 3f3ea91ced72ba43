#!/usr/bin/env bash
# One 2-GPU gpurun call (≈ 3.5 GPU-minutes):  /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/validate_2gpu.sh'
# multi-GPU tests (in-process and, new, multi-process over CUDA IPC), the multicast probe, and the N = 2 bench line.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multigpu_gpu.py tests/test_multiprocess_gpu.py -q -x 2>&1 | tail -6
timeout 60 tools/_build/multicast_probe 2>&1 | tee gpurun_out/multicast_probe.jsonl
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/n2.err | grep '^{' > gpurun_out/bench_n2.json
python - <<'PY'
import json
r = json.load(open("gpurun_out/bench_n2.json"))
print("N=2:", round(r["value"], 1), r["unit"], "| ms/step", round(r["ms_per_step"], 3), "| pull GB/s/dir", round(r["roofline"]["achieved"], 1),
      "| e2e", round(r["e2e"]["value"], 1), "| launches", r["gpu_launches"])
PY
