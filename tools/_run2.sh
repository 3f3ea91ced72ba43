mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multigpu_gpu.py -q -x 2>&1 | tail -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/bench_n2_dma.json
python -c "
import json; r=json.load(open('gpurun_out/bench_n2_dma.json')); print(r['value'], r['ms_per_step'], r['roofline']['achieved'], r['roofline']['frac'], r['e2e'].get('value'), r['gpu_launches'], r['clocks'])"
