#!/bin/bash
# N = 2: the driver's launch line for bench.py (weak-scaling headline + slab block), then the reference arm under torchrun
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 \
   > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2_bench_n2.json'))
    print('N=2 value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
    for s in d.get('slab', []):
        print('  slab', json.dumps(s)[:600])
except Exception as e:
    print('FAILED', e)
PY
tail -5 gpurun_out/r2_bench_n2.err
