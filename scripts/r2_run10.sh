#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fastfft.py tests/test_gpu_functionals.py tests/test_gpu_system.py -q -m gpu -x > gpurun_out/r2j_pytest.log 2>&1
tail -4 gpurun_out/r2j_pytest.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step', round(d['ms_per_step'], 4), 'E', d['config_detail'].get('energy_Ha'), 'e2e', d['e2e']['value'])
    for k in d['roofline']['kernels']:
        print('    %-70s %8.1f us' % (k['stage'], 1e3 * k['ms_per_eval']))
    print('   also', json.dumps(d.get('also')))
except Exception as e:
    print('FAILED', e)
PY
}
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-denopt > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; summ gpurun_out/r2j_bench.json
PAD_BENCH_GRID=512 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-denopt > gpurun_out/r2j_bench512.json 2> gpurun_out/r2j_bench512.err; summ gpurun_out/r2j_bench512.json
tail -3 gpurun_out/r2j_bench512.err
