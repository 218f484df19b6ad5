#!/bin/bash
# round 2, run 2: plain path with the fused mid kernel (default) vs the pipelined kernels; parity suites for the new code
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fastfft.py tests/test_gpu_spectral.py -x -q -m gpu > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -6 gpurun_out/r2b_pytest.log
for cfg in "0 0 0" "1 8 4"; do
  set -- $cfg
  PAD_PIPE=$1 PAD_PIPE_LPI=$2 PAD_PIPE_TPI=$3 timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-denopt \
     > gpurun_out/r2b_bench_pipe$1.json 2> gpurun_out/r2b_bench_pipe$1.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2b_bench_pipe$1.json'))
    print('pipe=$1', 'ms/step', round(d['ms_per_step'], 4), 'E', d.get('config_detail', d['config']).get('energy_Ha'), 'e2e', d['e2e']['value'])
    for k in d['roofline']['kernels']:
        print('    %-70s %8.1f us' % (k['stage'], 1e3 * k['ms_per_eval']))
    print('   also', json.dumps(d.get('also')))
except Exception as e:
    print('pipe=$1 FAILED', e)
PY
done
tail -3 gpurun_out/r2b_bench_pipe0.err
