#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pme.py tests/test_gpu_geometry.py tests/test_gpu_ionion.py -q -m gpu > gpurun_out/r2i_pytest.log 2>&1
tail -5 gpurun_out/r2i_pytest.log
for i in 1 2 3; do timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-denopt 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'clk samples', d['clocks'].get('samples'))"; done
