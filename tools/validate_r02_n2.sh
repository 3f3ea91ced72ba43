#!/usr/bin/env bash
# the round's last 2-GPU call:  /usr/local/graft/bin/gpurun --gpus 2 --timeout 420 -- 'bash tools/validate_r02_n2.sh'
# N = 2 bench line as the driver launches it (ring over NVLink, one runtime over both GPUs for e2e, worker-process chain), then
# the tests that need two GPUs.
set -u
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/bench_n2_final.err | grep '^{' > gpurun_out/bench_n2_final.json
python - <<'PY'
import json
r = json.load(open("gpurun_out/bench_n2_final.json"))
print("N=2:", round(r["value"], 1), r["unit"], "| ms/step", round(r["ms_per_step"], 3), "| GB/s/dir", round(r["roofline"]["achieved"], 1),
      "| share", round(r["roofline"]["kernel_share_of_step"], 3), "| dma", r["roofline"]["dma_copies"],
      "| e2e", round(r["e2e"]["value"], 1), round(r["e2e"]["ms_per_step"], 2), "| api dev", round(r["api"]["device_resident"]["ms_per_step"], 2),
      "| parity", r["parity"]["engines"], r["parity"]["bit_exact"], r["parity"]["corruption_detected"], "| clocks", r["clocks"]["sm_mhz"], r["clocks"]["reasons"])
for c in r["extra"]["configs"]:
    print("  extra:", json.dumps(c)[:330])
PY
tail -3 gpurun_out/bench_n2_final.err
timeout 150 python -m pytest tests/test_multigpu_gpu.py tests/test_multiprocess_gpu.py tests/test_grpc_gpu.py tests/test_differential_gpu.py -m gpu -q > gpurun_out/pytest_2gpu_final.log 2>&1; tail -4 gpurun_out/pytest_2gpu_final.log
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
