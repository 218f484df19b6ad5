mkdir -p gpurun_out
python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 1500 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('evals/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e'], d['clocks'], 'frac', round(d['roofline']['frac'],3), d.get('cpu_baseline'))
for k in d['roofline']['kernels']: print('   %-45s x%-3d %8.1f us  %s GB/s' % (k['stage'], k['launches_per_eval'], k['ms_per_eval']*1e3, round(k['GBps']) if k['GBps'] else None))
PY
