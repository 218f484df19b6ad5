#!/bin/bash
# N = 8: the driver's launch line for bench.py (weak-scaling headline + slab block)
mkdir -p gpurun_out
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
   > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f'gpurun_out/bench_n{n}.json'))
    print('N', n, 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
    for s in d.get('slab', []):
        print('  slab', {k: s.get(k) for k in ('workload', 'ms_per_eval', 'hbm_frac_per_gpu', 'nvlink_GBps_each_way', 'exchange', 'error', 'ms', 'ms_pme_order_8')})
except Exception as e:
    print('FAILED', e)
PY
tail -5 gpurun_out/bench_n$N.err
