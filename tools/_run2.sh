mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python -m pytest tests/test_multigpu_gpu.py -q -x -k "pull or corruption" 2>&1 | tail -5
for cfg in "auto 0 6" "tma 0 6" "dma 32 4" "dma 16 8" "dma 64 6"; do set -- $cfg
  timeout 200 $T bench.py --gpus 2 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --engine $1 --dma-last-mib $2 --dma-growth $3 2>/dev/null | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('$1 last=$2 growth=$3', round(r['value'],1), round(r['ms_per_step'],3), round(r['roofline']['achieved'],1), r['roofline'].get('dma_copies'), r['gpu_launches'])
" ; done
