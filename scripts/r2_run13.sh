#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fastfft.py tests/test_gpu_parallel.py tests/test_gpu_system.py -q -m gpu -x > gpurun_out/r2m_pytest.log 2>&1
tail -6 gpurun_out/r2m_pytest.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step', round(d['ms_per_step'], 4), d['launch_detail'], 'e2e', d['e2e']['value'])
    for k in d['roofline']['kernels']:
        print('    %-70s %8.1f us' % (k['stage'], 1e3 * k['ms_per_eval']))
    print('   also', json.dumps(d.get('also')))
    print('   denopt', json.dumps(d.get('density_optimization'))[:1500])
except Exception as e:
    print('FAILED', e)
PY
}
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; summ gpurun_out/r2m_bench.json
tail -3 gpurun_out/r2m_bench.err
