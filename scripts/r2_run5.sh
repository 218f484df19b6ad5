#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ionion.py tests/test_gpu_pme.py tests/test_gpu_ions.py tests/test_gpu_system.py tests/test_gpu_geometry.py tests/test_gpu_functionals.py -q -m gpu > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -25 gpurun_out/r2e_pytest.log
bash scripts/r2_ncu2.sh
PAD_BENCH_GRID=512 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-denopt > gpurun_out/r2e_bench512.json 2> gpurun_out/r2e_bench512.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2e_bench512.json'))
    print('512^3 ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'])
    for k in d['roofline']['kernels']:
        print('    %-70s %8.1f us' % (k['stage'], 1e3 * k['ms_per_eval']))
except Exception as e:
    print('512 FAILED', e)
PY
tail -3 gpurun_out/r2e_bench512.err
