#!/bin/bash
# round 2, run 4: folded kernel table A/B, fused term list with the unfused mid, full GPU suite, ncu --set full with source counters
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -6 gpurun_out/r2d_pytest.log
for cfg in "1" "0"; do
  PAD_FOLD_TABLE=$cfg timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-denopt \
     > gpurun_out/r2d_bench_fold$cfg.json 2> gpurun_out/r2d_bench_fold$cfg.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2d_bench_fold$cfg.json'))
    print('fold=$cfg', 'ms/step', round(d['ms_per_step'], 4), 'E', d.get('config_detail', d['config']).get('energy_Ha'), 'e2e', d['e2e']['value'])
    for k in d['roofline']['kernels']:
        print('    %-70s %8.1f us' % (k['stage'], 1e3 * k['ms_per_eval']))
    print('   also', json.dumps(d.get('also')))
except Exception as e:
    print('fold=$cfg FAILED', e)
PY
done
bash scripts/r2_ncu1.sh
