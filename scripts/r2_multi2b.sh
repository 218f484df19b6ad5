#!/bin/bash
mkdir -p gpurun_out
for k in 20 50; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps $k --warmup 3 --no-slab \
   2> gpurun_out/r2_bench_n2b.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2 steps', d['steps'], 'ms/step', d['ms_per_step'], d['launch_detail'], d['clocks'])"
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-denopt 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 ms/step', d['ms_per_step'], d['launch_detail'])"
