mkdir -p gpurun_out
N=$1
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/weak_n$N.json 2> gpurun_out/weak_n$N.err
tail -3 gpurun_out/weak_n$N.err; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/weak_n$N.json') if l.startswith('{')][-1])
print('WEAK n_gpus', d['n_gpus'], 'evals/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'])
PY
for grid in 512 1024; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 5 --warmup 3 --slab-grid $grid > gpurun_out/slab_${grid}_n$N.json 2> gpurun_out/slab_${grid}_n$N.err
tail -3 gpurun_out/slab_${grid}_n$N.err; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/slab_${grid}_n$N.json') if l.startswith('{')][-1])
    print('SLAB grid', $grid, 'n_gpus', d['n_gpus'], 'evals/s', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'E', d['config']['energy_Ha'], 'nvlink GB/s', round(d['roofline']['nvlink_GBps_each_way'],1))
except Exception as e: print('slab', $grid, 'failed', e)
PY
done
