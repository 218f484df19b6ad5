mkdir -p gpurun_out
for zg in 0 2 4; do
  PAD_FAST_FFT=1 PAD_ZGROUP=$zg python bench.py --steps 100 --no-cpu-baseline > gpurun_out/bench_fast_zg$zg.json 2> gpurun_out/bench_fast_zg$zg.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_fast_zg$zg.json'))
    print('zgroup', $zg, 'evals/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'])
    for k in d['roofline']['kernels']: print('   %-45s x%-3d %8.1f us  %s GB/s' % (k['stage'], k['launches_per_eval'], k['ms_per_eval']*1e3, round(k['GBps']) if k['GBps'] else None))
except Exception as e:
    print('zgroup', $zg, 'failed', e); print(open('gpurun_out/bench_fast_zg$zg.err').read()[-2000:])
PY
done
