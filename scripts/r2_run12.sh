#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parallel.py -q -m gpu -x -k "term_list or fused" > gpurun_out/r2l_pytest.log 2>&1
tail -15 gpurun_out/r2l_pytest.log
for i in 1 2; do timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-denopt 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'clocks', d['clocks'])"; done
