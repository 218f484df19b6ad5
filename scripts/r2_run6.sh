#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fastfft.py tests/test_gpu_functionals.py -q -m gpu > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -8 gpurun_out/r2f_pytest.log
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-denopt > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2f_bench.json'))
    print('ms/step', round(d['ms_per_step'], 4), 'E', d.get('config_detail', d['config']).get('energy_Ha'), 'e2e', d['e2e']['value'])
    for k in d['roofline']['kernels']:
        print('    %-70s %8.1f us' % (k['stage'], 1e3 * k['ms_per_eval']))
    print('   also', json.dumps(d.get('also')))
except Exception as e:
    print('FAILED', e)
PY
timeout 600 python scripts/dbg_geom.py > gpurun_out/r2f_geom.log 2>&1
grep -E "<<<<|ok |Iter" gpurun_out/r2f_geom.log | head -20; grep -c call gpurun_out/r2f_geom.log
bash scripts/r2_ncu2.sh > /dev/null 2>&1
