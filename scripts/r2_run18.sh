#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fastfft.py tests/test_gpu_functionals.py -q -m gpu -x 2>&1 | tail -2
for y in 0 1; do
PAD_XONE=$y timeout 600 python bench.py --steps 50 --warmup 8 --no-cpu-baseline --no-denopt 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('xone $y: ms/step', round(d['ms_per_step'],4), d['also']['values'])
for k in d['roofline']['kernels']:
    if 'x-fwd' in k['stage']: print('    %-40s %8.1f us' % (k['stage'], 1e3*k['ms_per_eval']))"
done
